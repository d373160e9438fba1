"""world_size-2 gloo tests (CPU) of the multi-rank host logic: the pair-range partition, the
(B,L,72) column-summary exchange protocol and the distance gather.  The arithmetic on each
rank is done by the oracle (test infrastructure); what is under test is that summing the
72-float summaries of disjoint pair ranges and concatenating the per-rank distances
reproduces the unsharded result, and that Phyloformer._gather reassembles ragged ranges."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests._util import GOLDEN


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pf_oracle
        from phyloformer_b200 import sharding
        from phyloformer_b200.model import Phyloformer
        torch.set_num_threads(2)
        ck = torch.load(os.path.join(GOLDEN, "ckpt_pf.pt"), map_location="cpu")
        w = pf_oracle.strip_prefix(ck["state_dict"])
        idx = pf_oracle.synth_msa(7, 21, seed=3, B=2)        # P = 21 -> ranges of 11 and 10
        lo, hi = sharding.pair_range(7, rank, world)
        calls = []

        def red(t):
            assert t.shape == (2, 21, 72)
            calls.append(1)
            dist.all_reduce(t)
            return t

        local = pf_oracle.forward_idx(w, idx, torch.float64, pair_lo=lo, pair_hi=hi, reduce_fn=red)
        assert len(calls) == pf_oracle.NB                     # one exchange per block
        m = Phyloformer()
        m.shard_pairs()
        full = m._gather(local.float(), 7)
        q.put((rank, full.numpy()))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "".join(traceback.format_exception(e))))
    finally:
        dist.destroy_process_group()


def test_pair_sharded_forward_with_gloo():
    from oracle import pf_oracle
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    outs = dict(q.get(timeout=300) for _ in range(2))
    [p.join(60) for p in procs]
    for r in range(2):
        assert not isinstance(outs[r], str), outs[r]
    assert all(p.exitcode == 0 for p in procs)
    ck = torch.load(os.path.join(GOLDEN, "ckpt_pf.pt"), map_location="cpu")
    w = pf_oracle.strip_prefix(ck["state_dict"])
    ref = pf_oracle.forward_idx(w, pf_oracle.synth_msa(7, 21, seed=3, B=2), torch.float64).numpy()
    for r in range(2):
        assert outs[r].shape == ref.shape
        assert np.max(np.abs(outs[r] - ref) / np.abs(ref)) < 1e-6
    assert np.array_equal(outs[0], outs[1])
