"""CPU-side pins of the two numerical devices of the tensor-core FFN kernel (no GPU needed):
the exp2-polynomial GELU, with its coefficients read from the kernel source, and the three-term
bf16 split product."""
import math
import os
import re

import numpy as np
import torch

from tests._util import ROOT


def _gelu_coeffs():
    src = open(os.path.join(ROOT, "phyloformer_b200", "csrc", "pf_ffn_ws.cuh")).read()
    body = src[src.index("__device__ __forceinline__ u64 gelu_fast2"):]
    body = body[:body.index("return fma2(pk2(-t0, -t1)")]
    vals = [float(v) for v in re.findall(r"pk2\((-?[0-9.]+e[+-][0-9]+)f,", body)]
    assert len(vals) == 7, vals            # degree-6 polynomial, highest power first
    return vals


def test_gelu_polynomial_matches_erf_gelu():
    """max(h,0) - t*exp2(p(t)), t = min(|h|,10), evaluated in fp32 like the kernel (Horner with
    fused multiply-adds emulated in fp64 then rounded), against the exact erf GELU."""
    c = _gelu_coeffs()
    h = np.concatenate([np.linspace(-12, 12, 200001), np.array([0.0, -0.0, 1e-8, -1e-8, 30.0, -30.0])]).astype(np.float32)
    t = np.minimum(np.abs(h), np.float32(10.0)).astype(np.float32)
    p = np.full_like(t, np.float32(c[0]))
    for ck in c[1:]:
        p = (p.astype(np.float64) * t.astype(np.float64) + np.float32(ck).astype(np.float64)).astype(np.float32)   # one rounding per FMA
    e = np.exp2(p.astype(np.float64)).astype(np.float32)
    g = (np.maximum(h, 0).astype(np.float64) - t.astype(np.float64) * e.astype(np.float64)).astype(np.float32)
    exact = np.array([0.5 * v * (1.0 + math.erf(v / math.sqrt(2.0))) for v in h.astype(np.float64)])
    err = np.abs(g.astype(np.float64) - exact)
    assert err.max() < 1e-6, float(err.max())          # 4.8e-7 abs
    assert g[-2] == np.float32(30.0) and abs(float(g[-1])) < 1e-20   # clamped tails: t E(t) < 1e-20 at t = 10


def test_bf16x3_split_product_error():
    """hi*hi + hi*lo + lo*hi with bf16 parts and fp32 accumulation (what the three MMA passes
    compute) reproduces an fp32 dot product to ~2^-16 relative; a single bf16 pass does not."""
    g = torch.Generator().manual_seed(0)
    a = torch.randn(512, 256, generator=g)
    w = torch.randn(256, 64, generator=g) * 0.1

    def split(x):
        hi = x.to(torch.bfloat16).to(torch.float32)
        lo = (x - hi).to(torch.bfloat16).to(torch.float32)
        return hi, lo

    ah, al = split(a)
    wh, wl = split(w)
    exact = a.double() @ w.double()
    three = (ah.double() @ wh.double()) + (ah.double() @ wl.double()) + (al.double() @ wh.double())
    one = ah.double() @ wh.double()
    scale = (a.abs().double() @ w.abs().double())
    e3 = ((three - exact).abs() / scale).max().item()
    e1 = ((one - exact).abs() / scale).max().item()
    assert e3 < 2.0 ** -15 and e1 > 20 * e3, (e3, e1)


def test_bias_fold_through_the_redundant_layernorm_channel():
    """WS_FOLD63 (pf_ffn_ws.cuh / pf_pack_ffn_tc): LayerNorm output sums to zero over its 64 channels, so
    W u + b == W' u' with W'[:, c] = W[:, c] - W[:, 63] (c < 63), W'[:, 63] = b and u' = u with channel 63 set to 1.
    Exact in exact arithmetic; with bf16 hi/lo operands (three-term product) the error stays at the split's level."""
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2048, 64, generator=g, dtype=torch.float64) * 3 + torch.randn(2048, 1, generator=g, dtype=torch.float64)
    u = torch.nn.functional.layer_norm(x, (64,), None, None, 1e-5)
    assert u.sum(-1).abs().max() < 1e-12
    w = torch.randn(256, 64, generator=g, dtype=torch.float64) * 0.2
    b = torch.randn(256, generator=g, dtype=torch.float64)
    wf = w.clone()
    wf[:, :63] -= w[:, 63:64]
    wf[:, 63] = b
    uf = u.clone()
    uf[:, 63] = 1.0
    exact = u @ w.T + b
    assert (uf @ wf.T - exact).abs().max() < 1e-12

    def split(t):
        t32 = t.to(torch.float32)
        hi = t32.to(torch.bfloat16).to(torch.float32)
        lo = (t32 - hi).to(torch.bfloat16).to(torch.float32)
        return hi.double(), lo.double()

    def three(a, m):
        ah, al = split(a)
        mh, ml = split(m)
        return ah @ mh.T + ah @ ml.T + al @ mh.T

    scale = u.abs() @ w.abs().T + b.abs()
    e_plain = ((three(u, w) + b - exact).abs() / scale).max().item()
    e_fold = ((three(uf, wf) - exact).abs() / scale).max().item()
    assert e_plain < 2.0 ** -15 and e_fold < 2.0 ** -14, (e_plain, e_fold)     # the folded weights are ~sqrt(2) larger
