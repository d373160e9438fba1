"""pytest configuration: markers, path setup and shared fixtures.

`-m "not gpu"` tests run on CPU in the build container; `-m gpu` tests run on a B200
and call the CUDA path through the C-ABI (phyloformer_b200/libpf_sm100.so).
Neither set reads /root/reference: goldens live in tests/golden/.
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run under gpurun)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def pf_weights():
    from oracle import pf_oracle
    ck = torch.load(os.path.join(GOLDEN, "ckpt_pf.pt"), map_location="cpu")
    return pf_oracle.strip_prefix(ck["state_dict"])


@pytest.fixture(scope="session")
def pf_indel_weights():
    from oracle import pf_oracle
    ck = torch.load(os.path.join(GOLDEN, "ckpt_pf_indel.pt"), map_location="cpu")
    return pf_oracle.strip_prefix(ck["state_dict"])


@pytest.fixture(scope="session")
def ref_cases():
    return dict(np.load(os.path.join(GOLDEN, "ref_cases.npz")))


@pytest.fixture(scope="session")
def ref_testdata():
    return dict(np.load(os.path.join(GOLDEN, "ref_testdata_pf.npz")))


def rel_err(a, b):
    """max and mean relative error of a against b (b = reference)."""
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    r = np.abs(a - b) / np.maximum(np.abs(b), 1e-30)
    return float(r.max()), float(r.mean())
