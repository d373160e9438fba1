#!/usr/bin/env python
"""Fit and check the fast GELU used by the tcgen05 FFN epilogues (pf_ffn_ws.cuh: gelu_fast2,
pf_ffn_tc.cuh: gelu_fast).

    GELU(h) = max(h,0) - t E(t),   t = min(|h|, 10),   E(t) = 0.5 erfc(t/sqrt 2) = exp2(-Q(t))

Q is a degree-6 polynomial, minimax-fitted (Lawson iterations on Chebyshev nodes of [0,6]) with
weight t E(t), i.e. minimising the absolute error of GELU itself.  The check evaluates the
kernel's exact fp32 operation order (Horner FMAs, exp2, one FMA)."""
import numpy as np
from numpy.polynomial import chebyshev as C
from scipy.special import erf, erfc

f32 = np.float32


def fit(deg=6, tmax=6.0, n=8000, iters=120):
    k = np.arange(n)
    u = np.cos(np.pi * (k + 0.5) / n)
    t = (u + 1) / 2 * tmax
    Q = -np.log2(0.5 * erfc(t / np.sqrt(2)))
    w = np.maximum(t * 0.5 * erfc(t / np.sqrt(2)), 1e-4)
    V = C.chebvander(u, deg)
    lw = np.ones(n)
    for _ in range(iters):
        W = (w * lw)[:, None]
        c = np.linalg.lstsq(V * W, Q * w * lw, rcond=None)[0]
        r = np.abs((V @ c - Q) * w)
        lw *= 0.3 + r / r.max()
        lw /= lw.mean()
    pu = C.cheb2poly(c)
    base, acc, cur = np.array([-1.0, 2 / tmax]), np.zeros(deg + 1), np.array([1.0])
    for ci in pu:
        acc[: len(cur)] += ci * cur
        cur = np.convolve(cur, base)
    return acc  # Q(t) = sum acc[i] t^i


def gelu_fast(h, coef):
    h = h.astype(f32)
    t = np.minimum(np.abs(h), f32(10.0))
    p = np.full_like(t, f32(-coef[-1]))
    for k in coef[-2::-1]:
        p = (p * t + f32(-k)).astype(f32)
    e = np.exp2(p.astype(np.float64)).astype(f32)
    return (np.maximum(h, f32(0)) - (t * e).astype(f32)).astype(f32)


if __name__ == "__main__":
    co = fit()
    print("Q coefficients t^0..t^6:", ", ".join("%.10e" % x for x in co))
    h = np.linspace(-60, 60, 4800001)
    ex = 0.5 * h * (1 + erf(h / np.sqrt(2)))
    err = np.abs(gelu_fast(h, co).astype(np.float64) - ex)
    print("max abs err %.3e at h=%.3f, rms %.3e" % (err.max(), h[err.argmax()], np.sqrt((err ** 2).mean())))
