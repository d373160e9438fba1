#!/usr/bin/env python
"""Run under torchrun on N GPUs: the pair-sharded forward (NVLink peer exchange and NCCL) must
reproduce the unsharded forward on every rank.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29533 tools/check_sharded.py
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    from phyloformer.model import Phyloformer
    ck = torch.load(os.path.join(ROOT, "tests", "golden", "ckpt_pf.pt"), map_location="cpu")
    m = Phyloformer(**ck["hyper_parameters"], precision="bf16x3")
    m.load_state_dict({k.replace("model.", ""): v for k, v in ck["state_dict"].items() if k != "model.seq2pair"},
                      strict=False)
    m = m.to("cuda").eval()
    g = torch.Generator().manual_seed(11)
    worst = 0.0
    for B, n, L in ((1, 64, 300), (2, 33, 130), (1, 9, 1000)):
        idx = torch.randint(0, 21, (B, n, L), dtype=torch.uint8, generator=g).cuda()
        m.unshard()
        full = m.forward_idx(idx, squeeze=False)
        m.shard_pairs(exchange="p2p")
        a = m.forward_idx(idx, squeeze=False)
        a2 = m.forward_idx(idx, squeeze=False)
        m.shard_pairs(exchange="nccl")
        b = m.forward_idx(idx, squeeze=False)
        m.check_device_error()
        assert torch.equal(a, a2), "p2p exchange is not deterministic"
        rel = ((a - full).abs() / full.abs()).max().item()
        reln = ((b - full).abs() / full.abs()).max().item()
        worst = max(worst, rel, reln)
        assert rel < 1e-4 and reln < 1e-4, (B, n, L, rel, reln)
        ref = [torch.empty_like(a) for _ in range(world)]
        dist.all_gather(ref, a)
        assert all(torch.equal(r, a) for r in ref), "ranks disagree"
    # a weight update re-creates the native handle: the peer buffers are re-bound (flags zeroed, epochs restart)
    m.shard_pairs(exchange="p2p")
    before = m.forward_idx(idx, squeeze=False)
    with torch.no_grad():
        m.pwFNN[0].bias.add_(0.0)
    after = m.forward_idx(idx, squeeze=False)
    after2 = m.forward_idx(idx, squeeze=False)
    m.check_device_error()
    assert torch.equal(before, after) and torch.equal(after, after2), "re-bound peer exchange differs"
    if rank == 0:
        print(f"sharded == unsharded on {world} ranks (p2p and nccl), worst max-rel {worst:.2e}; exchange impl "
              f"{os.environ.get('PF_EXCH_IMPL', 'fused')}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
