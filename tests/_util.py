"""Helpers shared by the tests (error metrics, fixture loading)."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def rel_err(a, b):
    """(max, mean) relative error of a against the reference b."""
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    assert a.shape == b.shape, (a.shape, b.shape)
    r = np.abs(a - b) / np.maximum(np.abs(b), 1e-30)
    return float(r.max()), float(r.mean())


def list_stems():
    return sorted(f[:-3] for f in os.listdir(os.path.join(GOLDEN, "msas")) if f.endswith(".fa"))
