#!/usr/bin/env python
"""Accuracy of the fast GELU used by the tcgen05 FFN epilogue (pf_ffn_tc.cuh: gelu_fast).

GELU(h) = max(h,0) - |h| * 0.5 erfc(|h|/sqrt 2),  0.5 erfc(z) ~ 1 / (c p(z))^16 with the
Abramowitz & Stegun 7.1.28 polynomial, c = 2^(1/16) and the 1/sqrt2 folded into the
coefficients.  Evaluated here with the same fp32 operation order as the kernel."""
import numpy as np
from scipy.special import erf

f32 = np.float32
a = np.array([0.0705230784, 0.0422820123, 0.0092705272, 0.0001520143, 0.0002765672, 0.0000430638])
c = 2 ** (1 / 16)
coef = [c] + [c * a[i] * (2 ** -0.5) ** (i + 1) for i in range(6)]


def gelu_fast(h):
    h = h.astype(f32)
    t = np.abs(h)
    p = np.full_like(t, f32(coef[6]))
    for k in coef[5::-1]:
        p = (p * t + f32(k)).astype(f32)
    for _ in range(4):
        p = (p * p).astype(f32)
    with np.errstate(over="ignore", divide="ignore"):
        r = (f32(1) / p).astype(f32)
    return (np.maximum(h, f32(0)) - np.abs((h * r).astype(f32))).astype(f32)


if __name__ == "__main__":
    print("coefficients:", ", ".join("%.10ef" % x for x in coef))
    h = np.linspace(-12, 12, 6000001)
    ex = 0.5 * h * (1 + erf(h / np.sqrt(2)))
    err = np.abs(gelu_fast(h).astype(np.float64) - ex)
    print("max abs err %.3e at h=%.4f, rms %.3e" % (err.max(), h[err.argmax()], np.sqrt((err ** 2).mean())))
