"""Generate the full-size value fixtures for BASELINE configs 3 and 5 (run in the build container):

    python tests/golden/make_fullsize.py 200x1000     # ~6 min on 8 vCPU,  10 GB RSS
    python tests/golden/make_fullsize.py 500x500      # ~20 min,           33 GB RSS

The unmodified reference cannot produce these (it needs ~70 / ~218 GB at these shapes and raises
for n > 200, reference phyloformer/model.py:24-28), so they come from the pair-chunked fp64 oracle
(oracle/pf_oracle.forward_streaming), which tests/test_oracle_golden.py pins (a) against the
monolithic oracle and (b), through it, against outputs of the reference itself at the sizes the
reference can run.  The inputs are exactly bench.py's synthetic MSAs for these workloads
(pf_oracle.synth_msa(n, L, seed=1337 + n, kind="tree")), so the same file checks the benchmarked
computation.  Output: tests/golden/oracle_fullsize_<n>x<L>.npz  {dist: float64 (P,), seed, n, L}.
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import pf_oracle  # noqa: E402


def main(shape):
    n, L = (int(v) for v in shape.lower().split("x"))
    seed = 1337 + n
    ck = torch.load(os.path.join(HERE, "ckpt_pf.pt"), map_location="cpu")
    w = pf_oracle.strip_prefix(ck["state_dict"])
    idx = pf_oracle.synth_msa(n, L, seed=seed, kind="tree")
    t0 = time.time()
    d = pf_oracle.forward_streaming(w, idx, torch.float64, chunk=256,
                                    progress=lambda s: print(f"[{time.time() - t0:7.1f}s] {s}", flush=True))
    out = os.path.join(HERE, f"oracle_fullsize_{n}x{L}.npz")
    np.savez_compressed(out, dist=d[0].numpy(), seed=seed, n=n, L=L)
    print(f"wrote {out}: {d.shape[1]} distances, min {float(d.min()):.3e} max {float(d.max()):.3e}, "
          f"{time.time() - t0:.0f} s")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "200x1000")
