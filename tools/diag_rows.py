#!/usr/bin/env python
"""GPU diagnostic: which rows/sites of the block-1 row-attention output differ between the FFMA row kernel
(PF_ROW_IMPL=tma) and the tcgen05 one (PF_ROW_IMPL=tc)?   python tools/diag_rows.py [n L]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(n, L, row, out):
    os.environ["PF_COL_IMPL"], os.environ["PF_ROW_IMPL"] = "cc", row
    import torch
    from oracle import pf_oracle
    from phyloformer.model import Phyloformer
    ck = torch.load(os.path.join(ROOT, "tests", "golden", "ckpt_pf.pt"), map_location="cpu")
    m = Phyloformer(**ck["hyper_parameters"], precision="bf16x3")
    m.load_state_dict({k.replace("model.", ""): v for k, v in ck["state_dict"].items() if k != "model.seq2pair"}, strict=False)
    m = m.to("cuda").eval()
    idx = pf_oracle.synth_msa(n, L, seed=1337 + n).cuda()
    a = m.debug_activation(idx, 4)          # after block 1's row attention
    torch.cuda.synchronize()
    m.check_device_error()
    torch.save(a[0].cpu(), out)


def main(n, L):
    for row in ("tma", "tc"):
        subprocess.run([sys.executable, __file__, "one", str(n), str(L), row, f"/tmp/act_{row}.pt"], check=True)
    import torch
    a, b = torch.load("/tmp/act_tma.pt"), torch.load("/tmp/act_tc.pt")     # (P, L, 64)
    scale = float(a.abs().max())
    err = (a - b).abs().amax(dim=2) / scale                              # (P, L)
    bad = (err > 1e-4)
    rows = bad.any(dim=1).nonzero().flatten().tolist()
    print(f"n={n} L={L}: scale {scale:.3f}, max err {float(err.max()):.3e}, {len(rows)} bad rows of {a.shape[0]}")
    for r in rows[:60]:
        sites = bad[r].nonzero().flatten()
        print(f"  row {r} (row % 148 = {r % 148}, ordinal {r // 148}): {len(sites)} bad sites, first {int(sites[0])} last {int(sites[-1])}, "
              f"max err {float(err[r].max()):.3e}, tiles {sorted(set((sites // 128).tolist()))}")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        worker(int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5])
    else:
        main(int(sys.argv[1]) if len(sys.argv) > 2 else 200, int(sys.argv[2]) if len(sys.argv) > 2 else 1000)
