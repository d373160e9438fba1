"""The drop-in boundary is a C ABI with no torch types in it: a plain C99 program
(tests/cabi/cabi_demo.c: include/pf_sm100.h + the CUDA runtime, nothing else) creates a handle
from raw weight buffers, runs pf_forward on an alignment and writes the distances.  They must be
bit-identical to what the Python host gets through ctypes, and match the oracle."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest
import torch

from oracle import pf_oracle
from tests._util import GOLDEN, ROOT, rel_err

pytestmark = pytest.mark.gpu


def test_plain_c_caller_matches_python_host(tmp_path, pf_weights):
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available on this box")
    from phyloformer.model import Phyloformer
    from phyloformer_b200.model import weight_names
    ck = torch.load(os.path.join(GOLDEN, "ckpt_pf.pt"), map_location="cpu")
    sd = {k.replace("model.", ""): v for k, v in ck["state_dict"].items() if k != "model.seq2pair"}
    names = weight_names(6)
    with open(tmp_path / "weights.bin", "wb") as f:
        f.write(struct.pack("<i", len(names)))
        for k in names:
            t = sd[k].detach().to(torch.float32).contiguous().numpy()
            f.write(struct.pack("<q", t.size))
            f.write(t.tobytes())
    idx = pf_oracle.synth_msa(17, 90, seed=5, B=2)
    with open(tmp_path / "msa.bin", "wb") as f:
        f.write(struct.pack("<iii", *idx.shape))
        f.write(idx.numpy().tobytes())
    libdir = os.path.join(ROOT, "phyloformer_b200")
    cuda_inc = "/usr/local/cuda/include"
    cudart_dir = next((d for d in ("/usr/local/cuda/lib64", os.path.join(os.path.dirname(torch.__file__), "..", "nvidia",
                                                                          "cuda_runtime", "lib"))
                       if os.path.isdir(d)), None)
    exe = str(tmp_path / "cabi_demo")
    cmd = [gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", cuda_inc,
           os.path.join(ROOT, "tests", "cabi", "cabi_demo.c"), "-o", exe, "-L", libdir, "-l:libpf_sm100.so",
           "-L", cudart_dir, "-lcudart", f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{cudart_dir}"]
    subprocess.run(cmd, check=True)
    r = subprocess.run([exe, str(tmp_path / "weights.bin"), str(tmp_path / "msa.bin"), str(tmp_path / "out.bin"), "1"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got = np.fromfile(tmp_path / "out.bin", dtype=np.float32).reshape(2, -1)

    m = Phyloformer(**ck["hyper_parameters"], precision="bf16x3")
    m.load_state_dict(sd, strict=False)
    m = m.to("cuda").eval()
    want = m.forward_idx(idx.cuda(), squeeze=False).cpu().numpy()
    assert np.array_equal(got, want)                      # same library, same inputs: bit-identical
    ref = pf_oracle.forward_idx(pf_weights, idx).numpy()
    assert rel_err(got, ref)[0] < 1e-3
