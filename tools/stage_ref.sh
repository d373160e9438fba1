#!/bin/bash
# Stage the third-party FastME binary the topology gate needs under baseline/_ref/ (git-ignored,
# but it travels to the GPU box with the gpurun snapshot).  Build container only.
set -e
mkdir -p baseline/_ref/bin
cp /root/reference/bin/bin_linux/fastme baseline/_ref/bin/fastme
chmod +x baseline/_ref/bin/fastme
echo staged baseline/_ref/bin/fastme
# the reference's own CLI, unmodified, for the drop-in test (tests/test_gpu_dropin.py)
cp /root/reference/infer_alns.py baseline/_ref/infer_alns.py
echo staged baseline/_ref/infer_alns.py
