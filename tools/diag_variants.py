#!/usr/bin/env python
"""GPU diagnostic: run one shape under a given PF_COL_IMPL / PF_ROW_IMPL pair in a fresh process and
report errors vs the fp64 oracle (or vs the all-FFMA attention build when the shape is too big).
    python tools/diag_variants.py                     # driver: loops over shapes x variants in subprocesses
    python tools/diag_variants.py one n L B col row   # worker"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(n, L, B, col, row):
    os.environ["PF_COL_IMPL"], os.environ["PF_ROW_IMPL"] = col, row
    os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
    import numpy as np
    import torch
    from oracle import pf_oracle
    from phyloformer.model import Phyloformer
    ck = torch.load(os.path.join(ROOT, "tests", "golden", "ckpt_pf.pt"), map_location="cpu")
    m = Phyloformer(**ck["hyper_parameters"], precision="bf16x3")
    m.load_state_dict({k.replace("model.", ""): v for k, v in ck["state_dict"].items() if k != "model.seq2pair"}, strict=False)
    m = m.to("cuda").eval()
    idx = pf_oracle.synth_msa(n, L, seed=1337 + n, B=B)
    d = m.forward_idx(idx.cuda(), squeeze=False)
    torch.cuda.synchronize()
    m.check_device_error()
    d = d.double().cpu().numpy()
    fx = os.path.join(ROOT, "tests", "golden", f"oracle_fullsize_{n}x{L}.npz")
    if B == 1 and os.path.exists(fx):
        ref = np.load(fx)["dist"][None]
    elif n * L * B <= 60000:
        ref = pf_oracle.forward_idx(pf_oracle.strip_prefix(ck["state_dict"]), idx, torch.float64).numpy()
    else:
        ref = None
    if ref is not None:
        rel = np.abs(d - ref) / np.abs(ref)
        i = int(np.argmax(rel))
        print(f"  n={n} L={L} B={B} col={col} row={row}: max-rel {rel.max():.3e} at d={ref.reshape(-1)[i]:.4e} (abs {abs(d.reshape(-1)[i]-ref.reshape(-1)[i]):.2e}), "
              f"mean-rel {rel.mean():.3e}, p99 {np.quantile(rel, 0.99):.3e}, min d {ref.min():.3e}", flush=True)
    else:
        print(f"  n={n} L={L} B={B} col={col} row={row}: ran, sum {d.sum():.6f}", flush=True)


def main():
    shapes = [(50, 500, 1), (100, 500, 1), (200, 1000, 1)]
    variants = [("tc", "tma"), ("cc", "tc"), ("tc", "tc")]
    for n, L, B in shapes:
        for col, row in variants:
            r = subprocess.run([sys.executable, __file__, "one", str(n), str(L), str(B), col, row], capture_output=True, text=True, timeout=120)
            out = r.stdout.strip()
            if r.returncode != 0:
                err = [ln for ln in r.stderr.splitlines() if "rror" in ln][:2]
                out += f"  n={n} L={L} B={B} col={col} row={row}: FAILED rc={r.returncode} {err}"
            print(out, flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        worker(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5], sys.argv[6])
    else:
        main()
