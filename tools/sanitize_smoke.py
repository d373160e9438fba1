"""Tiny forward in every precision mode for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import pf_oracle
from phyloformer_b200.model import Phyloformer

ck = torch.load("tests/golden/ckpt_pf.pt", map_location="cpu")
m = Phyloformer(**ck["hyper_parameters"])
m.load_state_dict({k.replace("model.", ""): v for k, v in ck["state_dict"].items() if k != "model.seq2pair"}, strict=False)
m = m.cuda().eval()
ref_w = pf_oracle.strip_prefix(ck["state_dict"])
for shape in ((7, 37, 1), (5, 16, 2), (4, 9, 19), (20, 300, 1)):   # (4,9,19): 171 sites -> several finalize CTAs with a ragged tail;
                                                               # (20,300,1): 190 pair-rows -> two rows per CTA in the persistent row kernel, 3 tiles per row
    n, L, B = shape
    idx = pf_oracle.synth_msa(n, L, seed=3, B=B)
    ref = pf_oracle.forward_idx(ref_w, idx).numpy()
    for prec in ("fp32", "bf16x3", "fp16", "bf16"):
        m.set_precision(prec)
        d = m.forward_idx(idx.cuda(), squeeze=False)
        torch.cuda.synchronize()
        if shape == (7, 37, 1):      # forward(x) with a soft (not one-hot) input: the FFMA block-0 row kernel instead of the residue-pair tables
            xs = (0.9 * pf_oracle.msa_to_onehot(idx) + 0.1 / 22).cuda()
            m(xs)
            torch.cuda.synchronize()
        m.check_device_error()
        err = float(((d.cpu().double().numpy() - ref) / ref).__abs__().max())
        print(shape, prec, "max-rel %.2e" % err)
print("done")
