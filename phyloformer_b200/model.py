"""Host-side mirror of the reference's `phyloformer.model.Phyloformer` for the inference path.

Same constructor, state-dict keys, `.to()/.eval()/load_state_dict()` behaviour and
`forward(x: (B,22,L,n) float) -> (P,) | (B,P)` contract as reference phyloformer/model.py:
109-201, so `infer_alns.py` and `models/*.ckpt` work unchanged.  The arithmetic is not done
by torch: `forward` hands device pointers to libpf_sm100.so (include/pf_sm100.h), which runs
hand-written sm_100a kernels on the current CUDA stream.  There is no CPU path.

Extras that the reference does not have:
  * `forward_idx(msa)`        (B,n,L)/(n,L) uint8 residue codes -> distances, skipping the one-hot
  * `distance_matrix(d, n)`   (B,P) -> (B,n,n) symmetric, on device (infer_alns.py:14-25)
  * `shard_pairs(group)`      pair-axis sharding over a torch.distributed process group
  * no `n <= 200` limit       (the reference's global SEQ2PAIR, model.py:24-28,39)
"""
import ctypes
import os
from typing import Optional

import torch
from torch import nn

from . import _cabi, sharding

__all__ = ["Phyloformer", "WEIGHT_ORDER", "weight_names"]


def weight_names(nb_blocks: int):
    """State-dict keys in the order pf_create expects them (reference layout)."""
    names = ["embedding_block.0.weight", "embedding_block.0.bias"]
    for i in range(nb_blocks):
        p = f"attention_blocks.{i}."
        for att in ("row_attention", "col_attention"):
            for proj in ("k_proj", "q_proj", "v_proj", "out_proj"):
                names += [f"{p}{att}.{proj}.weight", f"{p}{att}.{proj}.bias"]
        for norm in ("row_norm", "col_norm", "ffn_norm"):
            names += [f"{p}{norm}.weight", f"{p}{norm}.bias"]
        names += [f"{p}ffn.0.weight", f"{p}ffn.0.bias", f"{p}ffn.3.weight", f"{p}ffn.3.bias"]
    names += ["pwFNN.0.weight", "pwFNN.0.bias"]
    return names


WEIGHT_ORDER = weight_names(6)


class _AttentionParams(nn.Module):
    """Parameter container with the reference's ScaledLinearAttention names
    (attention.py:13-50,148-158): q/k project to one scalar per head."""

    def __init__(self, embed_dim: int, nb_heads: int):
        super().__init__()
        self.k_proj = nn.Linear(embed_dim, nb_heads)
        self.q_proj = nn.Linear(embed_dim, nb_heads)
        self.v_proj = nn.Linear(embed_dim, embed_dim)
        self.out_proj = nn.Linear(embed_dim, embed_dim)


class _BlockParams(nn.Module):
    """Parameter container for one axial block (model.py:45-85)."""

    def __init__(self, embed_dim: int, nb_heads: int, dropout: float):
        super().__init__()
        self.row_attention = _AttentionParams(embed_dim, nb_heads)
        self.col_attention = _AttentionParams(embed_dim, nb_heads)
        self.row_norm = nn.LayerNorm(embed_dim)
        self.col_norm = nn.LayerNorm(embed_dim)
        self.ffn_norm = nn.LayerNorm(embed_dim)
        # indices 0 and 3 carry the weights, as in the reference's nn.Sequential
        self.ffn = nn.Sequential(
            nn.Conv2d(embed_dim, 4 * embed_dim, kernel_size=1), nn.Dropout(dropout), nn.GELU(),
            nn.Conv2d(4 * embed_dim, embed_dim, kernel_size=1), nn.Dropout(dropout))


class Phyloformer(nn.Module):
    """B200-native Phyloformer (inference). See the module docstring."""

    def __init__(self, n_blocks: int = 6, n_heads: int = 4, h_dim: int = 64, dropout: float = 0.0,
                 n_seqs: int = 20, seq_len: int = 200, normalize: bool = True, heterodims: bool = False,
                 **kwargs):
        super().__init__()
        # The checkpoints store nb_blocks/nb_heads/embed_dim, which the reference constructor
        # swallows in **kwargs (model.py:112-123) and then builds the 6/4/64 default; accepting
        # both spellings gives the same network for every shipped checkpoint.
        n_blocks = int(kwargs.pop("nb_blocks", n_blocks))
        n_heads = int(kwargs.pop("nb_heads", n_heads))
        h_dim = int(kwargs.pop("embed_dim", h_dim))
        self.precision = kwargs.pop("precision", os.environ.get("PF_PRECISION", "bf16x3"))
        kwargs.pop("device", None)
        if n_heads != 4 or h_dim != 64:
            raise ValueError("phyloformer_b200 builds kernels for n_heads=4, h_dim=64 only "
                             f"(got {n_heads}, {h_dim})")
        self.nb_blocks, self.nb_heads, self.embed_dim = n_blocks, n_heads, h_dim
        self.dropout, self.normalize, self.heterodims = dropout, normalize, heterodims
        self.n_seqs, self.seq_len = n_seqs, seq_len

        self.embedding_block = nn.Sequential(nn.Conv2d(22, h_dim, kernel_size=1), nn.ReLU())
        self.attention_blocks = nn.ModuleList(
            [_BlockParams(h_dim, n_heads, dropout) for _ in range(n_blocks)])
        self.pwFNN = nn.Sequential(nn.Conv2d(h_dim, 1, kernel_size=1), nn.Dropout(dropout), nn.Softplus())

        self._handle = None
        self._handle_key = None
        self._handle_gen = 0  # bumped whenever the native handle is created or destroyed (graphed runners check it)
        self._ws = None
        self._shard = None  # (group, rank, world)
        self._exchange = "auto"
        self._peer = None
        self._peer_failed = False
        self._reduce_cb = None
        self.last_launches = 0
        # Pair-sharded forwards over the peer-memory exchange end with a (synchronising) read of the device error flag: a
        # peer that never publishes makes the bounded wait time out, and the result of that call is then invalid.  Set
        # to False to keep the forward asynchronous and call check_device_error() yourself (bench.py does).
        self.check_peer_errors = True

    # ------------------------------------------------------------------ native handle
    def _weights_key(self):
        sd = dict(self.named_parameters())
        return tuple((sd[k].data_ptr(), sd[k]._version) for k in weight_names(self.nb_blocks))

    def _ensure_handle(self, device):
        lib = _cabi.load()
        key = (device, self._weights_key())
        if self._handle is not None and key == self._handle_key:
            return lib
        self._release()
        if self.precision not in _cabi.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_cabi.PRECISIONS)}, got {self.precision!r}")
        sd = dict(self.named_parameters())
        tensors = []
        for k in weight_names(self.nb_blocks):
            t = sd[k].detach()
            if t.device != device or t.dtype != torch.float32 or not t.is_contiguous():
                t = t.to(device=device, dtype=torch.float32).contiguous()
            tensors.append(t)
        ptrs = (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
        cfg = _cabi.PfCfg(self.nb_blocks, self.nb_heads, self.embed_dim, 4, _cabi.PRECISIONS[self.precision])
        handle = ctypes.c_void_p()
        with torch.cuda.device(device):
            _cabi.check(lib.pf_create(ctypes.byref(handle), ctypes.byref(cfg), ptrs, len(tensors)), "pf_create")
        self._handle, self._handle_key = handle, key
        self._handle_gen += 1
        return lib

    def _release(self):
        if self._handle is not None:
            _cabi.load().pf_destroy(self._handle)
            self._handle = None
            self._handle_key = None
            self._handle_gen += 1

    def __del__(self):
        try:
            self._release()
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass

    def set_precision(self, precision: str):
        """'fp32' (FFMA, ~1e-6 of the reference), 'bf16x3' (tcgen05, 3-term split), or the fast modes 'fp16' / 'bf16' (no activation split)."""
        if precision not in _cabi.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_cabi.PRECISIONS)}")
        self.precision = precision
        if self._handle is not None:
            _cabi.check(_cabi.load().pf_set_precision(self._handle, _cabi.PRECISIONS[precision]), "pf_set_precision")

    # ------------------------------------------------------------------ sharding
    def shard_pairs(self, group=None, exchange: str = "auto"):
        """Shard the pair axis over `group` (default: the world group). Every rank must call
        forward with the same input; every rank gets the full result.

        exchange: how the per-block (B,L,72) column summaries are summed across ranks
          "nccl"  torch.distributed.all_reduce through the C ABI's reduce callback
          "p2p"   the library's own kernels over NVLink peer memory (torch symmetric memory
                  provides the mapped buffers; no collective library on the data path)
          "auto"  p2p when the backend is NCCL and symmetric memory can be set up, else nccl"""
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        if exchange not in ("auto", "nccl", "p2p"):
            raise ValueError("exchange must be 'auto', 'nccl' or 'p2p'")
        world = dist.get_world_size(group)
        self._shard = (group, dist.get_rank(group), world) if world > 1 else None
        self._exchange = exchange
        self._peer = None          # (symmetric tensor, handle, slot_floats, handle generation it is bound to)
        self._peer_failed = False
        return self

    def unshard(self):
        """Back to one rank computing every pair (undoes shard_pairs)."""
        self._shard = None
        return self

    def _setup_peer_exchange(self, lib, need_floats, device):
        """(Re)allocate the symmetric exchange buffer; collective over the shard group."""
        import torch.distributed as dist
        group, rank, world = self._shard
        if self._exchange == "nccl" or self._peer_failed:
            return False
        if self._exchange == "auto" and dist.get_backend(group) != "nccl":
            return False
        if self._peer is not None and self._peer[2] >= need_floats:
            if self._peer[3] != self._handle_gen:   # the handle was re-created (weights changed): bind the same buffers again
                buf, hdl, slot, _ = self._peer
                torch.cuda.synchronize(device)
                dist.barrier(group)       # nobody still reads the old epochs' flags or slots
                buf.zero_()               # the new handle counts its exchange epochs from zero again
                torch.cuda.synchronize(device)
                dist.barrier(group)
                ptrs = (ctypes.c_void_p * world)(*[int(p) for p in hdl.buffer_ptrs])
                _cabi.check(lib.pf_set_peer_exchange(self._handle, rank, world, ptrs, slot), "pf_set_peer_exchange")
                self._peer = (buf, hdl, slot, self._handle_gen)
            return True
        try:
            import torch.distributed._symmetric_memory as symm_mem
            slot = int(need_floats)
            nbytes = lib.pf_peer_exchange_bytes(slot)
            buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
            buf.zero_()
            torch.cuda.synchronize(device)
            hdl = symm_mem.rendezvous(buf, group if group is not None else dist.group.WORLD)
            dist.barrier(group)           # every rank has zeroed its flags before anyone publishes
            ptrs = (ctypes.c_void_p * world)(*[int(p) for p in hdl.buffer_ptrs])
            _cabi.check(lib.pf_set_peer_exchange(self._handle, rank, world, ptrs, slot), "pf_set_peer_exchange")
            self._peer = (buf, hdl, slot, self._handle_gen)
            return True
        except Exception as e:  # noqa: BLE001
            if self._exchange == "p2p":
                raise
            import warnings
            warnings.warn(f"peer-memory exchange unavailable ({e!r}); using NCCL all-reduce")
            self._peer_failed = True
            return False

    def _make_reduce(self, ws):
        import torch.distributed as dist
        group = self._shard[0]
        base = ws.data_ptr()

        def _reduce(_user, buf, count, _stream):
            try:
                off = buf - base
                view = ws[off:off + 4 * count].view(torch.float32)
                dist.all_reduce(view, op=dist.ReduceOp.SUM, group=group)
                return 0
            except Exception:  # noqa: BLE001  (must not unwind through C)
                import traceback
                traceback.print_exc()
                return 1

        return _cabi.REDUCE_FN(_reduce)

    # ------------------------------------------------------------------ forward
    def _workspace(self, nbytes, device):
        if self._ws is None or self._ws.device != device or self._ws.numel() < nbytes:
            self._ws = None  # free before growing
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        return self._ws

    def _run(self, idx, x, flag, B, n, L, debug_stages=None):
        device = idx.device
        if device.type != "cuda":
            raise RuntimeError("phyloformer_b200 runs on a B200 (sm_100a) only; there is no CPU fallback")
        lib = self._ensure_handle(device)
        P = sharding.n_pairs(n)
        if n < 2:
            raise ValueError("need at least 2 sequences")
        if self._shard is not None:
            _, rank, world = self._shard
            lo, hi = sharding.pair_range(n, rank, world)
        else:
            lo, hi = 0, P
        with torch.cuda.device(device):
            nbytes = lib.pf_workspace_bytes(self._handle, B, n, L, lo, hi)
            if nbytes == 0 and hi > lo:
                raise _cabi.PfError("pf_workspace_bytes rejected the shape")
            stream = torch.cuda.current_stream(device).cuda_stream
            dist_out = torch.empty((B, hi - lo), dtype=torch.float32, device=device)
            act = None
            if hi > lo:
                ws = self._workspace(max(nbytes, 256), device)
                cb = _cabi.NULL_REDUCE
                if self._shard is not None:
                    if not self._setup_peer_exchange(lib, B * L * _cabi.PF_COLSUM_FLOATS, device):
                        _cabi.check(lib.pf_set_peer_exchange(self._handle, 0, 0, None, 0), "pf_set_peer_exchange")
                    cb = self._make_reduce(ws)   # used only when the peer exchange is off
                    self._reduce_cb = cb  # keep alive for the duration of the call
                xp = x.data_ptr() if x is not None else None
                fp = flag.data_ptr() if flag is not None else None
                if debug_stages is None:
                    rc = lib.pf_forward(self._handle, idx.data_ptr(), xp, fp, B, n, L, lo, hi, dist_out.data_ptr(),
                                        ws.data_ptr(), ws.numel(), stream, cb, None)
                else:
                    act = torch.empty((B, hi - lo, L, 64), dtype=torch.float32, device=device)
                    rc = lib.pf_forward_debug(self._handle, idx.data_ptr(), xp, fp, B, n, L, lo, hi,
                                              dist_out.data_ptr(), ws.data_ptr(), ws.numel(), stream, cb, None,
                                              int(debug_stages), act.data_ptr())
                _cabi.check(rc, "pf_forward")
                self.last_launches = lib.pf_last_launch_count(self._handle)
            elif self._shard is not None:
                raise RuntimeError("more ranks than pairs: use fewer ranks or batch sharding")
        if debug_stages is not None:
            return act
        if self._shard is not None:
            dist_out = self._gather(dist_out, n)
            if self.check_peer_errors and self._peer is not None:
                self.check_device_error()
        return dist_out

    def _gather(self, local, n):
        import torch.distributed as dist
        group, rank, world = self._shard
        ranges = sharding.all_ranges(n, world)
        width = max(hi - lo for lo, hi in ranges)
        B = local.shape[0]
        pad = torch.zeros((B, width), dtype=local.dtype, device=local.device)
        pad[:, : local.shape[1]] = local
        out = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(out, pad, group=group)
        return torch.cat([out[r][:, : hi - lo] for r, (lo, hi) in enumerate(ranges)], dim=1)

    def forward(self, input):
        """input: (B, 22, L, n) float tensor on a CUDA device (reference model.py:166-187).
        Returns torch.squeeze of the (B, P) distances, like the reference."""
        if input.dim() != 4 or input.shape[1] != 22:
            raise ValueError(f"expected input of shape (batch, 22, seq_len, n_seqs), got {tuple(input.shape)}")
        if not input.is_cuda:
            raise RuntimeError("phyloformer_b200 runs on a B200 (sm_100a) only; there is no CPU fallback "
                               "(move the model and the input to 'cuda')")
        B, _, L, n = input.shape
        x = input.detach()
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.to(torch.float32).contiguous()
        self.n_seqs, self.seq_len = n, L
        lib = _cabi.load()
        with torch.cuda.device(x.device):
            idx = torch.empty((B, n, L), dtype=torch.uint8, device=x.device)
            flag = torch.empty(1, dtype=torch.int32, device=x.device)
            stream = torch.cuda.current_stream(x.device).cuda_stream
            _cabi.check(lib.pf_onehot_to_idx(x.data_ptr(), B, L, n, idx.data_ptr(), flag.data_ptr(), stream),
                        "pf_onehot_to_idx")
        out = self._run(idx, x, flag, B, n, L)
        return torch.squeeze(out).to(input.dtype)

    def forward_idx(self, msa: torch.Tensor, squeeze: bool = True):
        """msa: (B, n, L) or (n, L) uint8 residue codes (ALPHABET order) on a CUDA device."""
        if msa.dim() == 2:
            msa = msa[None]
        if msa.dtype != torch.uint8:
            raise TypeError("forward_idx expects uint8 residue codes")
        if not msa.is_cuda:
            raise RuntimeError("phyloformer_b200 runs on a B200 (sm_100a) only; there is no CPU fallback")
        msa = msa.contiguous()
        B, n, L = msa.shape
        out = self._run(msa, None, None, B, n, L)
        return torch.squeeze(out) if squeeze else out

    def make_graphed(self, example_idx: torch.Tensor):
        """Capture forward_idx for one input shape into a CUDA graph (the ~32 kernel launches of a
        forward become one graph launch: small alignments are launch-bound otherwise).

        Returns `run(idx) -> (B, P) distances`; `idx` must have the shape/dtype of `example_idx`.
        The result tensor is reused between calls (clone it to keep it).  Not available while the
        pair axis is sharded (the exchange is a host-driven collective)."""
        if self._shard is not None:
            raise RuntimeError("make_graphed is not supported with shard_pairs()")
        ex = example_idx[None] if example_idx.dim() == 2 else example_idx
        if ex.dtype != torch.uint8 or not ex.is_cuda:
            raise TypeError("make_graphed expects a CUDA uint8 tensor of residue codes")
        static_in = ex.contiguous().clone()
        self.forward_idx(static_in, squeeze=False)          # warm-up: handle, lazy init
        torch.cuda.synchronize(static_in.device)
        # The graph bakes in raw pointers to the workspace and to the handle's device buffers.  It therefore gets
        # a workspace of its own (held by the closure; eager calls keep using / regrowing self._ws) and remembers
        # the handle generation and precision it was captured with: a weight change, .to() or set_precision()
        # afterwards makes run() raise instead of replaying against freed memory.
        B, n, L = static_in.shape
        lib = _cabi.load()
        nbytes = lib.pf_workspace_bytes(self._handle, B, n, L, 0, sharding.n_pairs(n))
        private_ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=static_in.device)
        eager_ws, self._ws = self._ws, private_ws
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = self.forward_idx(static_in, squeeze=False)
        finally:
            self._ws = eager_ws
        gen, prec = self._handle_gen, self.precision

        def run(idx: torch.Tensor) -> torch.Tensor:
            idx = idx[None] if idx.dim() == 2 else idx
            if idx.shape != static_in.shape or idx.dtype != torch.uint8:
                raise ValueError(f"graphed forward expects uint8 {tuple(static_in.shape)}, got {idx.dtype} {tuple(idx.shape)}")
            if self._handle_gen != gen or self.precision != prec or self._handle_key != (static_in.device, self._weights_key()):
                raise RuntimeError("the model changed (weights, device or precision) after make_graphed(): capture again")
            static_in.copy_(idx, non_blocking=True)
            graph.replay()
            return static_out

        run.workspace = private_ws  # owned by the runner
        run.graph = graph  # keep alive / introspection
        return run

    def debug_activation(self, input_or_idx, n_stages: int):
        """Test hook: activation (B,Pl,L,64) after `n_stages` sub-blocks (0 = pair embedding)."""
        t = input_or_idx
        if t.dtype == torch.uint8:
            t = t[None] if t.dim() == 2 else t
            B, n, L = t.shape
            return self._run(t.contiguous(), None, None, B, n, L, debug_stages=n_stages)
        B, _, L, n = t.shape
        x = t.detach().to(torch.float32).contiguous()
        lib = _cabi.load()
        idx = torch.empty((B, n, L), dtype=torch.uint8, device=x.device)
        flag = torch.empty(1, dtype=torch.int32, device=x.device)
        with torch.cuda.device(x.device):
            stream = torch.cuda.current_stream(x.device).cuda_stream
            _cabi.check(lib.pf_onehot_to_idx(x.data_ptr(), B, L, n, idx.data_ptr(), flag.data_ptr(), stream),
                        "pf_onehot_to_idx")
        return self._run(idx, x, flag, B, n, L, debug_stages=n_stages)

    def distance_matrix(self, d: torch.Tensor, n: int):
        """(P,) or (B,P) distances -> (B,n,n) symmetric matrices with zero diagonal, on device."""
        lib = _cabi.load()
        d2 = d.reshape(-1, sharding.n_pairs(n)).to(torch.float32).contiguous()
        B = d2.shape[0]
        mat = torch.empty((B, n, n), dtype=torch.float32, device=d2.device)
        with torch.cuda.device(d2.device):
            stream = torch.cuda.current_stream(d2.device).cuda_stream
            _cabi.check(lib.pf_dist_to_matrix(d2.data_ptr(), B, n, mat.data_ptr(), stream), "pf_dist_to_matrix")
        return mat

    def debug_tc_dump(self, idx: torch.Tensor):
        """Test hook: run one forward and return the raw tcgen05 accumulators of the first
        128-token tile of the LAST block's FFN kernel as a (128, 320) tensor."""
        lib = _cabi.load()
        idx = idx[None] if idx.dim() == 2 else idx
        self.forward_idx(idx)  # make sure the handle exists
        # [128][320] accumulators, followed by the per-CTA role timers of the profiling build
        # (3 roles x 8 floats per CTA, one CTA per SM): see pf_debug_set_dump in pf_sm100.h
        n_sm = torch.cuda.get_device_properties(idx.device).multi_processor_count
        flat = torch.zeros(128 * 320 + n_sm * 24, dtype=torch.float32, device=idx.device)
        dump = flat[:128 * 320].view(128, 320)
        _cabi.check(lib.pf_debug_set_dump(self._handle, dump.data_ptr()), "pf_debug_set_dump")
        try:
            self.forward_idx(idx)
            torch.cuda.synchronize(idx.device)
        finally:
            _cabi.check(lib.pf_debug_set_dump(self._handle, None), "pf_debug_set_dump")
        return dump

    def check_device_error(self):
        """Synchronise and raise if a kernel reported a device-side pipeline timeout."""
        if self._handle is not None:
            _cabi.check(_cabi.load().pf_device_error(self._handle), "device check")

    def profile_enable(self, on: bool = True):
        """Bracket every kernel of the following forwards with CUDA events (bench.py)."""
        if self._handle is None:
            raise RuntimeError("run one forward first")
        _cabi.check(_cabi.load().pf_profile_enable(self._handle, 1 if on else 0), "pf_profile_enable")

    def profile_read(self):
        """{kernel class: (total ms, launches)} since the last read. Synchronises."""
        ms = (ctypes.c_float * len(_cabi.KERNEL_CLASSES))()
        cnt = (ctypes.c_int32 * len(_cabi.KERNEL_CLASSES))()
        _cabi.check(_cabi.load().pf_profile_read(self._handle, ms, cnt), "pf_profile_read")
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(_cabi.KERNEL_CLASSES)}

    def _set_seq2pair(self, n_seqs: int):
        """Kept for API compatibility (model.py:189-201): pair indices are closed-form on the
        device, there is no seq2pair matrix to rebuild."""
        self.n_seqs = n_seqs
        self.n_pairs = sharding.n_pairs(n_seqs)
