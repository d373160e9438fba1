#!/usr/bin/env python
"""CPU emulation of reduced-precision operand formats for the two FFN GEMMs (no GPU needed).

Runs the fp64 oracle graph on the 20 test alignments with the FFN's GEMM operands rounded the way a
candidate tensor-core mode would round them, and reports the distance error against the committed
reference outputs -- the question a new `precision` mode has to answer before a kernel is written
(SURVEY.md 7.4.1 did this once for whole-model modes; this tool separates the two GEMMs).

    python tools/prec_emulate.py [mode ...]

Operand formats:  x3 = bf16 hi + bf16 lo (16 mantissa bits),  f16 = one fp16 rounding,
f16x2 = fp16 hi + fp16 lo,  bf16 = one bf16 rounding,  exact = fp64.
A mode is  <A1>/<W1>/<H>/<W2>,  e.g. the shipped parity mode is  x3/x3/x3/x3  (lo.lo dropped);
the suffix +fold63 moves the first bias into the GEMM through the redundant LayerNorm channel.
Test infrastructure: imports oracle/ (allowed for tools and tests only).
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pf_oracle as O  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def rnd(x, fmt):
    if fmt == "exact":
        return x, None
    if fmt == "bf16":
        return x.to(torch.bfloat16).to(x.dtype), None
    if fmt == "f16":
        return x.to(torch.float32).to(torch.float16).to(x.dtype), None
    if fmt == "x3":
        x32 = x.to(torch.float32)
        hi = x32.to(torch.bfloat16).to(torch.float32)
        lo = (x32 - hi).to(torch.bfloat16).to(torch.float32)
        return hi.to(x.dtype), lo.to(x.dtype)
    if fmt == "f16x2":
        x32 = x.to(torch.float32)
        hi = x32.to(torch.float16).to(torch.float32)
        lo = (x32 - hi).to(torch.float16).to(torch.float32)
        return hi.to(x.dtype), lo.to(x.dtype)
    raise ValueError(fmt)


def qlinear(a, w, fa, fw):
    """a @ w.T with both operands rounded; hi/lo formats drop the lo.lo term like the kernels do."""
    ah, al = rnd(a, fa)
    wh, wl = rnd(w, fw)
    out = F.linear(ah, wh)
    if al is not None:
        out = out + F.linear(al, wh)
    if wl is not None:
        out = out + F.linear(ah, wl)
    return out


def forward(w, x, mode, dtype=torch.float64):
    fold63 = mode.endswith("+fold63")
    fa1, fw1, fh, fw2 = mode.replace("+fold63", "").split("/")
    x = x.to(dtype)
    B, C, L, n = x.shape
    We = w["embedding_block.0.weight"].to(dtype).reshape(O.D, O.N_CHAR)
    be = w["embedding_block.0.bias"].to(dtype)
    emb = F.relu(torch.einsum("bcln,dc->bnld", x, We) + be)
    pi, pj = O.pair_indices(n)
    h = emb[:, pi] + emb[:, pj]
    for b in range(O.NB):
        p = f"attention_blocks.{b}."
        h = h + O._attention(O._ln(h, w, p + "row_norm", dtype), w, p + "row_attention.", dtype)
        u = O._ln(h, w, p + "col_norm", dtype).transpose(1, 2)
        h = h + O._attention(u, w, p + "col_attention.", dtype).transpose(1, 2)
        # LayerNorm affine folded into W1 / b1 as the kernels do: the A operand is the plain normalised row
        g, bt = w[p + "ffn_norm.weight"].to(dtype), w[p + "ffn_norm.bias"].to(dtype)
        nrm = F.layer_norm(h, (O.D,), None, None, 1e-5)
        W1 = w[p + "ffn.0.weight"].to(dtype).reshape(4 * O.D, O.D)
        W2 = w[p + "ffn.3.weight"].to(dtype).reshape(O.D, 4 * O.D)
        b1 = w[p + "ffn.0.bias"].to(dtype) + W1 @ bt
        if fold63:
            # sum_c n_c = 0, so channel 63 is redundant: W1'[:, c] = W1[:, c] - W1[:, 63] (c < 63) gives the same product,
            # and the freed operand column carries the constant 1 against W1'[:, 63] = b1 (the bias rides on the GEMM)
            Wp = (W1 * g).clone()
            Wp[:, :63] -= Wp[:, 63:64].clone()
            Wp[:, 63] = b1
            a = nrm.clone()
            a[..., 63] = 1.0
            hid = F.gelu(qlinear(a, Wp, fa1, fw1))
        else:
            hid = F.gelu(qlinear(nrm, W1 * g, fa1, fw1) + b1)
        h = h + qlinear(hid, W2, fh, fw2) + w[p + "ffn.3.bias"].to(dtype)
    z = F.linear(h, w["pwFNN.0.weight"].to(dtype).reshape(1, O.D), w["pwFNN.0.bias"].to(dtype))
    return F.softplus(z[..., 0]).mean(dim=-1)


def main():
    modes = sys.argv[1:] or ["x3/x3/x3/x3", "x3/x3/f16/f16x2", "f16/f16x2/x3/x3", "f16/f16x2/f16/f16x2",
                             "x3/x3/f16x2/f16", "x3/x3/bf16/x3"]
    w = O.strip_prefix(torch.load(os.path.join(GOLDEN, "ckpt_pf.pt"), map_location="cpu")["state_dict"])
    ref = dict(np.load(os.path.join(GOLDEN, "ref_testdata_pf.npz")))
    from phyloformer_b200.data import load_alignment_idx
    xs = {}
    for stem in sorted(ref):
        idx, _ = load_alignment_idx(os.path.join(GOLDEN, "msas", stem + ".fa"))
        xs[stem] = O.msa_to_onehot(idx[None] if idx.dim() == 2 else idx)
    exact = {s: forward(w, xs[s], "exact/exact/exact/exact").numpy()[0] for s in xs}
    for mode in modes:
        mx, mxs, mean = 0.0, "", []
        for s in xs:
            d = forward(w, xs[s], mode).numpy()[0]
            r = np.abs(d - exact[s]) / np.abs(exact[s])
            if r.max() > mx:
                mx, mxs = float(r.max()), s
            mean.append(float(r.mean()))
        print(f"{mode:24s} max-rel {mx:.2e} ({mxs})  mean-rel {np.mean(mean):.2e}", flush=True)


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    main()
