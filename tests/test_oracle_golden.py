"""Pin the CPU oracle (oracle/pf_oracle.py) against outputs of the unmodified reference.

The reference has no tests of its own; tests/golden/*.npz were produced
by tests/golden/make_golden.py from /root/reference in the build container.
Tolerances: the reference's own fp32 noise (1 vs 8 threads) is 6.8e-6 max-rel, so the
oracle (fp64 and fp32) must sit within 5e-5 max-rel (measured: <= 2.3e-5, mean 1e-7) of the reference fp32 output.
"""
import os

import numpy as np
import pytest
import torch

from oracle import pf_oracle
from tests._util import GOLDEN, rel_err, list_stems

TOL = 5e-5


def test_pair_order_is_lexicographic():
    i, j = pf_oracle.pair_indices(5)
    assert list(zip(i.tolist(), j.tolist())) == [(a, b) for a in range(5) for b in range(a + 1, 5)]


def test_load_alignment_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_load_alignment.npz"))
    idx, ids = pf_oracle.parse_fasta_idx(os.path.join(golden_dir, "msas", "0_20_tips.fa"))
    assert list(g["ids"]) == ids
    oh = pf_oracle.msa_to_onehot(idx[None])[0]          # (22,L,n)
    assert oh.shape == g["aln"].shape
    assert np.array_equal(oh.numpy().astype(np.int8), g["aln"])


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_small_taps(pf_weights, golden_dir, dtype):
    g = np.load(os.path.join(golden_dir, "ref_small_taps.npz"))
    taps = {}
    d = pf_oracle.forward_idx(pf_weights, torch.from_numpy(g["idx"]), dtype, taps=taps)
    tol = 1e-4  # intermediates are compared in absolute terms scaled by their magnitude
    def close(a, b):
        a = a.double().numpy(); scale = np.abs(b).max() + 1e-30
        return np.abs(a - b).max() / scale
    assert close(taps["x0"], g["x0"]) < 1e-6
    prev = taps["x0"]
    for b in range(pf_oracle.NB):
        assert close(taps[f"b{b}.row"] - prev, g[f"b{b}.row_attn"]) < tol
        assert close(taps[f"b{b}.col"] - taps[f"b{b}.row"], g[f"b{b}.col_attn"]) < tol
        assert close(taps[f"b{b}.ffn"] - taps[f"b{b}.col"], g[f"b{b}.ffn_out"]) < tol
        assert close(taps[f"b{b}.ffn"], g[f"b{b}.out"]) < tol
        prev = taps[f"b{b}.ffn"]
    assert rel_err(d[0], g["dist"])[0] < TOL


@pytest.mark.parametrize("stem", list_stems()[:8])
def test_testdata_distances(pf_weights, ref_testdata, golden_dir, stem):
    idx, _ = pf_oracle.parse_fasta_idx(os.path.join(golden_dir, "msas", stem + ".fa"))
    d = pf_oracle.forward_idx(pf_weights, idx[None], torch.float64)[0]
    mx, mean = rel_err(d, ref_testdata[stem])
    assert mx < TOL and mean < 2e-6, (stem, mx, mean)


def test_cases(pf_weights, pf_indel_weights, ref_cases):
    c = ref_cases
    d = pf_oracle.forward(pf_weights, torch.from_numpy(c["soft_x"]), torch.float64)
    assert d.shape == (2, 10) and rel_err(d, c["soft_d"])[0] < TOL
    gi = torch.from_numpy(c["gap_idx"])
    assert rel_err(pf_oracle.forward_idx(pf_indel_weights, gi)[0], c["gap_d_indel"])[0] < TOL
    assert rel_err(pf_oracle.forward_idx(pf_weights, gi)[0], c["gap_d_pf"])[0] < TOL
    d = pf_oracle.forward_idx(pf_weights, torch.from_numpy(c["n2_idx"]))
    assert c["n2_d"].shape == () and pf_oracle.squeeze_like_reference(d).shape == ()
    assert rel_err(d, c["n2_d"])[0] < TOL
    d = pf_oracle.forward_idx(pf_weights, torch.from_numpy(c["l1_idx"]))
    assert rel_err(d[0], c["l1_d"])[0] < TOL
    d = pf_oracle.forward_idx(pf_weights, torch.from_numpy(c["b3_idx"]))
    assert d.shape == c["b3_d"].shape and rel_err(d, c["b3_d"])[0] < TOL
    d = pf_oracle.forward_idx(pf_weights, torch.from_numpy(c["u12_idx"]), torch.float32)
    assert rel_err(d[0], c["u12_d"])[0] < TOL
    d = pf_oracle.forward_idx(pf_weights, torch.from_numpy(c["dup_idx"]))
    assert rel_err(d[0], c["dup_d"])[0] < 2e-4   # near-zero distances between duplicates


def test_pair_sharded_oracle_matches_unsharded(pf_weights, ref_cases):
    """Pair-range sharding with a summed (B,L,72) column summary is exact algebra."""
    idx = torch.from_numpy(ref_cases["u12_idx"])
    full = pf_oracle.forward_idx(pf_weights, idx)
    P = full.shape[1]
    # emulate 3 shards in one process: run the shards in lock-step through a fake reducer
    import threading
    W = 3
    bounds = [(r * P // W, (r + 1) * P // W) for r in range(W)]
    barrier = threading.Barrier(W)
    slots, outs = [None] * W, [None] * W

    def run(r):
        def red(t):
            slots[r] = t
            barrier.wait()
            s = sum(slots[k] for k in range(W))
            barrier.wait()
            return s
        outs[r] = pf_oracle.forward_idx(pf_weights, idx, pair_lo=bounds[r][0], pair_hi=bounds[r][1], reduce_fn=red)

    ts = [threading.Thread(target=run, args=(r,)) for r in range(W)]
    [t.start() for t in ts]; [t.join() for t in ts]
    got = torch.cat(outs, dim=1)
    assert rel_err(got, full)[0] < 1e-10


def test_phylip_text_matches_reference(ref_testdata, golden_dir):
    for stem in ("0_20_tips", "3_50_tips"):
        _, ids = pf_oracle.parse_fasta_idx(os.path.join(golden_dir, "msas", stem + ".fa"))
        txt = pf_oracle.vec_to_phylip_text(torch.from_numpy(ref_testdata[stem]), ids)
        assert txt == open(os.path.join(golden_dir, f"ref_phylip_{stem}.phy")).read()


@pytest.mark.parametrize("chunk", [1, 7, 1000])
def test_streaming_oracle_equals_monolithic(pf_weights, ref_testdata, golden_dir, chunk):
    """forward_streaming (pair-chunked, two passes per block: what the full-size fixtures of
    BASELINE configs 3 and 5 were generated with) is the same graph as forward(): equal to fp64
    round-off for chunk sizes that do and do not divide P, batch > 1, and -- through the reference
    fixtures -- within the oracle's tolerance of the unmodified reference."""
    idx = pf_oracle.synth_msa(9, 23, seed=11, B=2)
    full = pf_oracle.forward_idx(pf_weights, idx, torch.float64)
    st = pf_oracle.forward_streaming(pf_weights, idx, torch.float64, chunk=chunk)
    assert st.shape == full.shape and rel_err(st, full)[0] < 1e-12
    st32 = pf_oracle.forward_streaming(pf_weights, idx, torch.float32, chunk=chunk)
    assert rel_err(st32, full)[0] < TOL
    if chunk == 7:
        stem = "0_20_tips"
        fa, _ = pf_oracle.parse_fasta_idx(os.path.join(golden_dir, "msas", stem + ".fa"))
        d = pf_oracle.forward_streaming(pf_weights, fa[None], torch.float64, chunk=37)[0]
        assert rel_err(d, ref_testdata[stem])[0] < TOL


@pytest.mark.parametrize("shape", ["200x1000", "500x500"])
def test_fullsize_fixture_is_consistent(pf_weights, golden_dir, shape):
    """The committed full-size fixtures (tests/golden/make_fullsize.py) belong to bench.py's inputs:
    recompute a few pairs' ROW-attention-free invariants cheaply -- here: the fixture has P entries,
    is finite and positive, and a 12-taxon sub-alignment of the same MSA run through the oracle
    correlates with the corresponding sub-matrix (column attention couples pairs, so the values
    differ slightly; identical ordering of near and far pairs is what is checked)."""
    path = os.path.join(golden_dir, f"oracle_fullsize_{shape}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    g = np.load(path)
    n, L = int(g["n"]), int(g["L"])
    d = g["dist"]
    assert d.shape == (n * (n - 1) // 2,) and np.isfinite(d).all() and (d > 0).all()
    idx = pf_oracle.synth_msa(n, L, seed=int(g["seed"]), kind="tree")
    sub = pf_oracle.forward_idx(pf_weights, idx[:, :12, :200], torch.float32)[0].numpy()
    iu = np.triu_indices(n, 1)
    D = np.zeros((n, n)); D[iu] = d
    big = D[:12, :12][np.triu_indices(12, 1)]
    assert np.corrcoef(np.log(sub), np.log(big))[0, 1] > 0.9
