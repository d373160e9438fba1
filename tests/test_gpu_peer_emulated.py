"""The peer-memory exchange kernel (k_col_exchange: chunk reduce -> per-(rank, CTA) flags -> P2P reads of every
rank's slot -> finalize) with world = 2 on ONE GPU.

On a multi-GPU box the same kernel runs over NVLink under tests/test_gpu_sharded.py, tools/check_sharded.py and
every `bench.py --gpus N` line.  A one-GPU box cannot do that through torch (symmetric-memory rendezvous refuses two
ranks on one device), so this test drives the C ABI directly: two native handles ("ranks") in one process, two
plain device buffers registered as each other's peers with pf_set_peer_exchange, each rank's forward of its own pair
range on its own stream.  The two streams run concurrently; every block's exchange kernels spin on each other's
flags exactly as they do across GPUs.  The alignments are short (few site groups -> few exchange CTAs), so the waiting
CTAs of one rank cannot keep the other rank's kernels off the SMs.
"""
import ctypes

import numpy as np
import pytest
import torch

from oracle import pf_oracle
from tests._util import rel_err
from tests.test_gpu_parity import make_model

pytestmark = pytest.mark.gpu


def _sharded_forward_two_ranks(models, bufs, slot, idx, exch_calls=1):
    from phyloformer_b200 import _cabi, sharding
    lib = _cabi.load()
    B, n, L = idx.shape
    P = sharding.n_pairs(n)
    dev = idx.device
    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
    outs, keep = [], []
    torch.cuda.synchronize(dev)
    for rank, (m, st) in enumerate(zip(models, streams)):
        lo, hi = sharding.pair_range(n, rank, 2)
        nbytes = lib.pf_workspace_bytes(m._handle, B, n, L, lo, hi)
        ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)
        out = torch.empty((B, hi - lo), dtype=torch.float32, device=dev)
        keep.append(ws)
        outs.append(out)
        _cabi.check(lib.pf_forward(m._handle, idx.data_ptr(), None, None, B, n, L, lo, hi, out.data_ptr(), ws.data_ptr(),
                                   ws.numel(), st.cuda_stream, _cabi.NULL_REDUCE, None), "pf_forward")
    torch.cuda.synchronize(dev)
    for m in models:
        m.check_device_error()
    return torch.cat(outs, dim=1)


@pytest.mark.parametrize("impl", ["fused", "split"])
def test_peer_exchange_two_ranks_on_one_gpu(pf_weights, impl, monkeypatch):
    from phyloformer_b200 import _cabi
    monkeypatch.setenv("PF_EXCH_IMPL", impl)
    lib = _cabi.load()
    ranks = [make_model("ckpt_pf.pt", "bf16x3") for _ in range(2)]
    single = make_model("ckpt_pf.pt", "bf16x3")
    warm = pf_oracle.synth_msa(4, 8, seed=1).cuda()
    for m in ranks + [single]:
        m.forward_idx(warm)                       # creates the native handles (under PF_EXCH_IMPL)
    cases = [(9, 40, 1, 21), (14, 33, 2, 22), (30, 24, 1, 23)]          # (n, L, B, seed): B * L <= 66 site groups
    slot = max(B * L for _, L, B, _ in cases) * _cabi.PF_COLSUM_FLOATS
    nbytes = lib.pf_peer_exchange_bytes(slot)
    bufs = [torch.zeros(nbytes, dtype=torch.uint8, device="cuda") for _ in range(2)]
    ptrs = (ctypes.c_void_p * 2)(*[b.data_ptr() for b in bufs])
    torch.cuda.synchronize()
    for rank, m in enumerate(ranks):
        _cabi.check(lib.pf_set_peer_exchange(m._handle, rank, 2, ptrs, slot), "pf_set_peer_exchange")
    try:
        for n, L, B, seed in cases:
            idx = pf_oracle.synth_msa(n, L, seed=seed, B=B).cuda()
            full = single.forward_idx(idx, squeeze=False)
            a = _sharded_forward_two_ranks(ranks, bufs, slot, idx)
            b = _sharded_forward_two_ranks(ranks, bufs, slot, idx)     # epochs advance, slots alternate
            assert torch.equal(a, b), (impl, n, L, B)
            assert a.shape == full.shape
            mx, _ = rel_err(a.cpu().numpy(), full.cpu().numpy())
            assert mx < 1e-4, (impl, n, L, B, mx)                      # only the order of the cross-pair sum differs
            ref = pf_oracle.forward_idx(pf_weights, idx.cpu(), torch.float64).numpy()
            assert rel_err(a.cpu().numpy(), ref)[0] < 1e-3
        # 6 blocks x 2 forwards x 3 cases: both ranks went through the same number of exchange epochs
        flags = [b[: 32 * 512 * 4].view(torch.int32).view(32, 512) for b in bufs]
        assert int(flags[0][1].max()) == int(flags[1][0].max()) == 36
    finally:
        for m in ranks:
            lib.pf_set_peer_exchange(m._handle, 0, 0, None, 0)
