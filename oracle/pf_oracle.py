"""CPU oracle for the Phyloformer inference hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch CPU restatement (torch CPU tensors, functional ops,
no nn.Module) of what the reference computes in

    /root/reference/phyloformer/model.py:166-187      Phyloformer.forward
    /root/reference/phyloformer/model.py:87-106       PhyloformerLayer.forward
    /root/reference/phyloformer/attention.py:160-197  ScaledLinearAttention.forward
    /root/reference/phyloformer/model.py:8-18         seq2pair (pair order)
    /root/reference/phyloformer/data.py:7-31          ALPHABET / load_alignment layout

It exists so that `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` have something to check the CUDA path
against (and to time beside it).  Nothing under `phyloformer_b200/` may import
it: the product path has no CPU fallback.

Parity pin: the reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the reference itself, generated in the build
container by `tests/golden/make_golden.py` (which imports /root/reference) and
committed under `tests/golden/`.  `tests/test_oracle_golden.py` checks this file
against those vectors (fp32: <= 2e-5 max-rel, the reference's own thread-count
noise is 7e-6; fp64 oracle vs fp32 reference: same bound).

All third-party arithmetic is PyTorch ATen (the reference pins torch 2.0.1, this
image has 2.11): LayerNorm eps=1e-5 biased variance, ELU alpha=1, GELU exact erf,
Softplus beta=1 threshold=20.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional, Tuple

import torch
import torch.nn.functional as F

ALPHABET = b"ARNDCQEGHILKMFPSTWYVX-"  # data.py:7
N_CHAR = 22
D = 64
H = 4
DH = 16
NB = 6
# per-site column summary exchanged between pair shards:
#   [0:4)  sum_p k~_h      [4:8) sum_p q~_h      [8:72) sum_p k~_h * v_{h,:}
COLSUM = 72

Weights = Dict[str, torch.Tensor]


def strip_prefix(state_dict: Dict[str, torch.Tensor]) -> Weights:
    """infer_alns.py:75-82: drop the 'model.' prefix and the stale seq2pair entry."""
    out = {}
    for k, v in state_dict.items():
        if k in ("model.seq2pair", "seq2pair"):
            continue
        out[k[6:] if k.startswith("model.") else k] = v.detach().cpu()
    return out


def pair_indices(n: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Lexicographic (i<j) pair order of model.py:13-17 == torch.triu_indices(n,n,1)."""
    ij = torch.triu_indices(n, n, offset=1)
    return ij[0], ij[1]


def msa_to_onehot(idx: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """(B,n,L) residue codes -> (B,22,L,n) one-hot, the layout of data.py:28-29 +
    infer_alns.py:112."""
    oh = F.one_hot(idx.long(), num_classes=N_CHAR)  # (B,n,L,22)
    return oh.permute(0, 3, 2, 1).to(dtype).contiguous()


def _attention(u, w, pfx, dtype, reduce_fn=None, n_total=None):
    """attention.py:160-197 on u of shape (B,R,N,64); attends along N.

    reduce_fn (optional) sums the un-normalised summaries over pair shards; it is
    only passed for column attention, where N is the (sharded) pair axis.
    """
    B, R, N, _ = u.shape
    k = F.linear(u, w[pfx + "k_proj.weight"].to(dtype), w[pfx + "k_proj.bias"].to(dtype))
    q = F.linear(u, w[pfx + "q_proj.weight"].to(dtype), w[pfx + "q_proj.bias"].to(dtype))
    v = F.linear(u, w[pfx + "v_proj.weight"].to(dtype), w[pfx + "v_proj.bias"].to(dtype))
    k = F.elu(k) + 1  # (B,R,N,H)
    q = F.elu(q) + 1
    v = v.view(B, R, N, H, DH)
    ksum = k.sum(dim=2)  # (B,R,H)
    qsum = q.sum(dim=2)
    kv = torch.einsum("brnh,brnhe->brhe", k, v)  # (B,R,H,DH)
    if reduce_fn is not None:
        packed = torch.cat([ksum, qsum, kv.reshape(B, R, H * DH)], dim=-1).contiguous()
        packed = reduce_fn(packed)  # (B,R,72)
        ksum, qsum = packed[..., 0:4], packed[..., 4:8]
        kv = packed[..., 8:].reshape(B, R, H, DH)
    n_att = N if n_total is None else n_total
    ctx = kv / ksum[..., None]  # == (k / k.sum).T @ v      attention.py:186-190
    qhat = q / (qsum / n_att)[:, :, None, :]  # q / q.mean   attention.py:183
    o = qhat[..., None] * ctx[:, :, None, :, :]  # (B,R,N,H,DH) attention.py:192
    o = o.reshape(B, R, N, D)
    return F.linear(o, w[pfx + "out_proj.weight"].to(dtype), w[pfx + "out_proj.bias"].to(dtype))


def _ln(x, w, name, dtype):
    return F.layer_norm(x, (D,), w[name + ".weight"].to(dtype), w[name + ".bias"].to(dtype), 1e-5)


def forward(
    w: Weights,
    x: torch.Tensor,
    dtype: torch.dtype = torch.float64,
    pair_lo: int = 0,
    pair_hi: Optional[int] = None,
    reduce_fn: Optional[Callable[[torch.Tensor], torch.Tensor]] = None,
    taps: Optional[dict] = None,
    nb: int = NB,
) -> torch.Tensor:
    """Distances for pairs [pair_lo, pair_hi) of the lexicographic pair list.

    x: (B,22,L,n) float (any values; one-hot in practice).  Returns (B, pair_hi-pair_lo)
    (no squeeze: callers apply model.py:185's torch.squeeze themselves).
    Internal layout is token-major (B,P,L,64) rather than the reference's (B,64,P,L);
    every op is applied along the same logical axes.
    `taps`, if given, is filled with intermediates for the golden comparison.
    """
    x = x.to(dtype)
    B, C, L, n = x.shape
    assert C == N_CHAR
    P = n * (n - 1) // 2
    pair_hi = P if pair_hi is None else pair_hi
    # embedding_block (model.py:138-143,173): per-sequence-site conv1x1 22->64 + ReLU
    We = w["embedding_block.0.weight"].to(dtype).reshape(D, N_CHAR)
    be = w["embedding_block.0.bias"].to(dtype)
    emb = F.relu(torch.einsum("bcln,dc->bnld", x, We) + be)  # (B,n,L,64)
    # pair representation (model.py:175): rows of seq2pair hold exactly two ones
    pi, pj = pair_indices(n)
    pi, pj = pi[pair_lo:pair_hi], pj[pair_lo:pair_hi]
    h = emb[:, pi] + emb[:, pj]  # (B,Pl,L,64)
    if taps is not None:
        taps["x0"] = h.clone()
    sharded = reduce_fn is not None
    for b in range(nb):
        p = f"attention_blocks.{b}."
        # row attention: rows = pairs, attends along sites          model.py:90-92
        h = h + _attention(_ln(h, w, p + "row_norm", dtype), w, p + "row_attention.", dtype)
        if taps is not None:
            taps[f"b{b}.row"] = h.clone()
        # column attention: rows = sites, attends along pairs       model.py:96-98
        u = _ln(h, w, p + "col_norm", dtype).transpose(1, 2)  # (B,L,Pl,64)
        a = _attention(u, w, p + "col_attention.", dtype,
                       reduce_fn=reduce_fn if sharded else None, n_total=P)
        h = h + a.transpose(1, 2)
        if taps is not None:
            taps[f"b{b}.col"] = h.clone()
        # FFN (model.py:69-85,102-104): conv1x1 64->256, GELU(erf), conv1x1 256->64
        u = _ln(h, w, p + "ffn_norm", dtype)
        W1 = w[p + "ffn.0.weight"].to(dtype).reshape(4 * D, D)
        W2 = w[p + "ffn.3.weight"].to(dtype).reshape(D, 4 * D)
        u = F.linear(F.gelu(F.linear(u, W1, w[p + "ffn.0.bias"].to(dtype))), W2,
                     w[p + "ffn.3.bias"].to(dtype))
        h = h + u
        if taps is not None:
            taps[f"b{b}.ffn"] = h.clone()
    # pwFNN (model.py:158-164,182) + site mean (model.py:185)
    z = F.linear(h, w["pwFNN.0.weight"].to(dtype).reshape(1, D), w["pwFNN.0.bias"].to(dtype))
    d = F.softplus(z[..., 0]).mean(dim=-1)  # (B,Pl)
    return d


def forward_streaming(
    w: Weights,
    idx: torch.Tensor,
    dtype: torch.dtype = torch.float64,
    chunk: int = 256,
    store_dtype: Optional[torch.dtype] = None,
    nb: int = NB,
    progress: Optional[Callable[[str], None]] = None,
) -> torch.Tensor:
    """forward_idx() for shapes whose intermediates do not fit in host memory at once
    (BASELINE configs 3 and 5: 200 x 1000 and 500 x 500; the reference itself needs ~70 / ~218 GB
    there and refuses n > 200, model.py:24-28).

    Same graph as forward() -- model.py:166-187, 87-106, attention.py:160-197 -- evaluated over
    chunks of `chunk` pairs with two passes per block, which is possible because the only
    cross-pair coupling is column attention's per-site sums (attention.py:183-190 with N = pairs):
      pass 1  row attention on the chunk (independent per pair), then the chunk's contribution to
              the column summaries sum_p k, sum_p q, sum_p k v^T of the block;
      pass 2  column attention output from the completed summaries, residual, FFN, residual.
    Only the (B,P,L,64) activation is kept (in `store_dtype`, default = dtype); arithmetic is in
    `dtype`.  idx: (B,n,L) residue codes.  Returns (B,P) in `dtype`.
    tests/test_oracle_golden.py checks it against forward() (chunk sizes that do not divide P).
    """
    store_dtype = dtype if store_dtype is None else store_dtype
    idx = idx.long()
    B, n, L = idx.shape
    P = n * (n - 1) // 2
    We = w["embedding_block.0.weight"].to(dtype).reshape(D, N_CHAR)
    be = w["embedding_block.0.bias"].to(dtype)
    table = F.relu(We.t() + be)          # (22,64): embedding of a one-hot residue (model.py:138-143)
    emb = table[idx]                     # (B,n,L,64)
    pi, pj = pair_indices(n)
    x = torch.empty((B, P, L, D), dtype=store_dtype)
    say = progress or (lambda s: None)

    def lin(u, name):
        return F.linear(u, w[name + ".weight"].to(dtype).reshape(-1, u.shape[-1]), w[name + ".bias"].to(dtype))

    for b in range(nb):
        p = f"attention_blocks.{b}."
        ksum = torch.zeros((B, L, H), dtype=dtype)
        qsum = torch.zeros((B, L, H), dtype=dtype)
        kv = torch.zeros((B, L, H, DH), dtype=dtype)
        for c0 in range(0, P, chunk):
            c1 = min(P, c0 + chunk)
            if b == 0:
                h = emb[:, pi[c0:c1]] + emb[:, pj[c0:c1]]                      # model.py:175
            else:
                h = x[:, c0:c1].to(dtype)
            h = h + _attention(_ln(h, w, p + "row_norm", dtype), w, p + "row_attention.", dtype)  # model.py:90-92
            u = _ln(h, w, p + "col_norm", dtype)                               # (B,Pc,L,64)
            k = F.elu(lin(u, p + "col_attention.k_proj")) + 1                  # (B,Pc,L,H)
            q = F.elu(lin(u, p + "col_attention.q_proj")) + 1
            v = lin(u, p + "col_attention.v_proj").view(B, c1 - c0, L, H, DH)
            ksum += k.sum(dim=1)
            qsum += q.sum(dim=1)
            kv += torch.einsum("bplh,bplhe->blhe", k, v)
            x[:, c0:c1] = h.to(store_dtype)
        ctx = kv / ksum[..., None]                                             # attention.py:186-190
        qmean = qsum / P                                                       # attention.py:183
        say(f"block {b}: summaries done")
        W1 = w[p + "ffn.0.weight"].to(dtype).reshape(4 * D, D)
        W2 = w[p + "ffn.3.weight"].to(dtype).reshape(D, 4 * D)
        for c0 in range(0, P, chunk):
            c1 = min(P, c0 + chunk)
            h = x[:, c0:c1].to(dtype)
            u = _ln(h, w, p + "col_norm", dtype)
            q = F.elu(lin(u, p + "col_attention.q_proj")) + 1
            qhat = q / qmean[:, None]                                          # (B,Pc,L,H)
            o = (qhat[..., None] * ctx[:, None]).reshape(B, c1 - c0, L, D)     # attention.py:192-193
            h = h + lin(o, p + "col_attention.out_proj")                       # model.py:97-98
            u = _ln(h, w, p + "ffn_norm", dtype)
            h = h + F.linear(F.gelu(F.linear(u, W1, w[p + "ffn.0.bias"].to(dtype))), W2,
                             w[p + "ffn.3.bias"].to(dtype))                    # model.py:102-104
            x[:, c0:c1] = h.to(store_dtype)
        say(f"block {b}: done")
    out = torch.empty((B, P), dtype=dtype)
    wh = w["pwFNN.0.weight"].to(dtype).reshape(1, D)
    for c0 in range(0, P, chunk):
        c1 = min(P, c0 + chunk)
        z = F.linear(x[:, c0:c1].to(dtype), wh, w["pwFNN.0.bias"].to(dtype))
        out[:, c0:c1] = F.softplus(z[..., 0]).mean(dim=-1)                     # model.py:182-185
    return out


def forward_idx(w: Weights, idx: torch.Tensor, dtype=torch.float64, **kw) -> torch.Tensor:
    """Same as forward() from (B,n,L) residue codes."""
    return forward(w, msa_to_onehot(idx, dtype), dtype, **kw)


def squeeze_like_reference(d: torch.Tensor) -> torch.Tensor:
    """model.py:185 applies torch.squeeze to the (B,P) result."""
    return torch.squeeze(d)


# ----------------------------------------------------------------------------------
# host-side pieces of the boundary that the reference keeps next to the model
# ----------------------------------------------------------------------------------
def parse_fasta_idx(path: str):
    """data.py:11-31 restated: returns ((n,L) uint8 residue codes, ids)."""
    lookup = {c: i for i, c in enumerate(ALPHABET)}
    seqs, ids = [], []
    with open(path, "rb") as fh:
        for line in fh:
            line = line.strip()
            if line.startswith(b">"):
                ids.append(line[1:].decode("utf8"))
                seqs.append([])
            else:
                seqs[-1].extend(lookup[c] for c in line)  # KeyError outside ALPHABET
    return torch.tensor(seqs, dtype=torch.uint8), ids


def vec_to_phylip_text(d: torch.Tensor, ids) -> str:
    """infer_alns.py:14-25 restated: symmetric matrix, '%.10f', one row per taxon."""
    n = len(ids)
    dm = torch.zeros((n, n), dtype=d.dtype)
    i, j = pair_indices(n)
    dm[i, j] = d
    dm = dm + dm.T
    s = f"{n}\n"
    for name, row in zip(ids, dm):
        s += f"{name} " + " ".join(f"{float(v):.10f}" for v in row) + "\n"
    return s


# ----------------------------------------------------------------------------------
# synthetic MSA generators (G1 tree-like, G2 gapped, G3 uniform)
# ----------------------------------------------------------------------------------
def synth_msa(n: int, L: int, seed: int = 1337, kind: str = "tree", B: int = 1) -> torch.Tensor:
    """(B,n,L) uint8 residue codes."""
    g = torch.Generator().manual_seed(seed)
    out = torch.empty((B, n, L), dtype=torch.uint8)
    for b in range(B):
        if kind == "uniform":
            out[b] = torch.randint(0, 20, (n, L), generator=g, dtype=torch.uint8)
            continue
        # random binary tree by sequential splitting: sequence k copies a random earlier
        # one and mutates a random fraction of sites; earlier one mutates a little too.
        seqs = [torch.randint(0, 20, (L,), generator=g, dtype=torch.uint8)]
        while len(seqs) < n:
            par = int(torch.randint(0, len(seqs), (1,), generator=g))
            for tgt in (par, None):
                rate = 0.01 + 0.29 * float(torch.rand(1, generator=g))
                base = seqs[par].clone()
                mask = torch.rand(L, generator=g) < rate
                sub = torch.randint(0, 20, (L,), generator=g, dtype=torch.uint8)
                base[mask] = sub[mask]
                if tgt is None:
                    seqs.append(base)
                else:
                    seqs[par] = base
        m = torch.stack(seqs[:n])
        if kind == "gapped":
            # geometric-length gap runs (code 21), about 5 % of cells
            n_runs = max(1, int(0.05 * n * L / 4))
            for _ in range(n_runs):
                r = int(torch.randint(0, n, (1,), generator=g))
                s = int(torch.randint(0, L, (1,), generator=g))
                u = max(float(torch.rand(1, generator=g)), 1e-9)
                ln = 1 + int(math.log(u) / math.log(0.75))
                m[r, s:s + ln] = 21
        out[b] = m
    return out
