#!/bin/bash
# Stage unmodified reference artefacts under baseline/_ref/ (git-ignored, but it travels to the GPU box
# with the gpurun snapshot).  Build container only (/root/reference is not on the GPU box).
set -e
mkdir -p baseline/_ref/bin
# third-party FastME binary for the topology gate
cp /root/reference/bin/bin_linux/fastme baseline/_ref/bin/fastme
chmod +x baseline/_ref/bin/fastme
echo staged baseline/_ref/bin/fastme
# the reference's own CLI, unmodified, for the drop-in test (tests/test_gpu_dropin.py)
cp /root/reference/infer_alns.py baseline/_ref/infer_alns.py
echo staged baseline/_ref/infer_alns.py
# the reference's model, unmodified, for bench.py's reference arm and cpu_baseline (kind "reference").
# The package directory gets a different name so that it cannot shadow this repository's `phyloformer`
# shim; model.py only uses a relative import of attention.py (reference phyloformer/model.py:5), and
# the reference's __init__.py is not copied because it pulls in data.py -> dendropy (not installed).
mkdir -p baseline/_ref/phyloformer_ref
cp /root/reference/phyloformer/model.py /root/reference/phyloformer/attention.py baseline/_ref/phyloformer_ref/
: > baseline/_ref/phyloformer_ref/__init__.py
echo staged baseline/_ref/phyloformer_ref/{model,attention}.py
