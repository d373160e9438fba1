"""Build recipe for libpf_sm100.so (nvcc, sm_100a only, in-tree).

    python -m phyloformer_b200.build          # or __graft_entry__.build()

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box
with the gpurun snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpf_sm100.so")
SOURCES = ["pf_api.cu"]
HEADERS = ["pf_common.cuh", "pf_kernels.cuh", "pf_ffn_tc.cuh", "pf_ffn_ws.cuh", "pf_attn_tc.cuh", "pf_bme.h", os.path.join("..", "..", "include", "pf_sm100.h")]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC)")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False, defines=(), out=None):
    """Compile csrc/*.cu for sm_100a into phyloformer_b200/libpf_sm100.so (or `out`, with extra -D
    `defines`, for A/B experiments: load it with PF_LIB=<path>)."""
    if out is None and not force and not needs_build():
        return LIB
    cmd = [
        _nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
        "-Xptxas", "-v" if verbose else "-O3", "--shared", "-Xcompiler", "-fPIC,-O2",
        "-o", out or LIB,
    ] + ["-D" + d for d in defines] + [os.path.join(CSRC, s) for s in SOURCES] + ["-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libpf_sm100.so")
    if verbose:
        sys.stderr.write(r.stdout + r.stderr)
    return out or LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
