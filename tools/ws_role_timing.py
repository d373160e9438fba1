#!/usr/bin/env python
"""Role timing of the warp-specialised FFN kernel (run under gpurun with PF_WS_PROF=1).
Prints, averaged over CTAs and per tile: total cycles and cycles spent in each barrier wait
for the producer (tid 0), the MMA issuer (tid 128) and one epilogue warp (tid 256)."""
import os, sys
os.environ["PF_WS_PROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import pf_oracle
from phyloformer_b200 import _cabi
from phyloformer_b200.model import Phyloformer

n, L = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (200, 1000)
ck = torch.load("tests/golden/ckpt_pf.pt", map_location="cpu")
m = Phyloformer(**ck["hyper_parameters"], precision="bf16x3")
m.load_state_dict({k.replace("model.", ""): v for k, v in ck["state_dict"].items() if k != "model.seq2pair"}, strict=False)
m = m.cuda().eval()
idx = pf_oracle.synth_msa(n, L, seed=1).cuda()
m.forward_idx(idx)
lib = _cabi.load()
dump = torch.zeros(128 * 320 + 148 * 24, dtype=torch.float32, device="cuda")
_cabi.check(lib.pf_debug_set_dump(m._handle, dump.data_ptr()), "dump")
m.forward_idx(idx)
torch.cuda.synchronize()
t = dump[128 * 320:].view(148, 3, 8).cpu()
P = n * (n - 1) // 2
tiles = ((L + 3) // 4) * ((P + 31) // 32) / 148.0      # 32 pairs x 4 sites per tile (WS_G x WS_S)
names = {0: ("producer", ["a1_free", "d2_free", "load+LN1", "wait_st", "-", "-", "-"]),
         1: ("mma", ["a1_full", "h_full[a]", "h_full[b]", "issue_g1", "issue_g2", "-", "-"]),
         2: ("epilogue", ["g1_done[a]", "g1_done[b]", "g2_done", "wait_ld", "wait_st", "e2", "-"])}
print(f"tiles per CTA {tiles:.1f}")
for r in range(3):
    tot = t[:, r, 0].mean().item() / tiles
    print(f"{names[r][0]:9s} total {tot:8.0f} cyc/tile  " + "  ".join(f"{nm} {t[:, r, 1 + k].mean().item() / tiles:7.0f}" for k, nm in enumerate(names[r][1])))
