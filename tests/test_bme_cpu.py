"""CPU tests of the host-side tree search pf_bme_tree (BIONJ + balanced NNI / SPR; csrc/pf_bme.h).

The known answers were written by the FastME 2.1.6.4 binary the reference ships (README.md:85-92 runs it on
every matrix): tests/golden/fastme_cases.{npz,json} (tests/golden/make_fastme_golden.py) and the reference
trees of the 20 test alignments, tests/golden/ref_trees_pf.json.  No device work is involved."""
import itertools
import json
import os
import re

import numpy as np
import pytest

from phyloformer_b200.bme import bme_tree, phylip_rounded
from phyloformer_b200.treecmp import bipartitions, parse_newick, patristic_distances, rf_distance
from tests._util import GOLDEN

MODES = {"none": dict(nni=False, spr=False), "nni": dict(nni=True, spr=False),
         "spr": dict(nni=False, spr=True), "both": dict(nni=True, spr=True)}


def _tri(v):
    n = int(round((1 + (1 + 8 * len(v)) ** 0.5) / 2))
    m = np.zeros((n, n))
    m[np.triu_indices(n, 1)] = v
    return m + m.T


def _aligned_patristic(a, b):
    na, da = patristic_distances(a)
    nb, db = patristic_distances(b)
    assert sorted(na) == sorted(nb)
    ia = [na.index(x) for x in sorted(na)]
    ib = [nb.index(x) for x in sorted(nb)]
    return np.asarray(da)[np.ix_(ia, ia)], np.asarray(db)[np.ix_(ib, ib)]


@pytest.fixture(scope="module")
def fastme_cases():
    mats = dict(np.load(os.path.join(GOLDEN, "fastme_cases.npz")))
    return mats, json.load(open(os.path.join(GOLDEN, "fastme_cases.json")))


@pytest.mark.parametrize("mode", list(MODES))
def test_same_tree_as_the_fastme_binary(fastme_cases, mode):
    """Every stage against FastME's own output: BIONJ alone, + NNI, + SPR, + both.  Same number of moves,
    same topology (strict RF = 0: these matrices have no ties), same branch lengths to the 8 printed digits,
    same tree lengths as FastME's log."""
    mats, want = fastme_cases
    assert len(mats) >= 40
    moves = 0
    for name, m in sorted(mats.items()):
        ids = [f"T{i + 1}" for i in range(m.shape[0])]
        ours, st = bme_tree(m, ids, return_stats=True, **MODES[mode])
        ref = want[name][mode]
        assert rf_distance(ours, ref["newick"]) == 0, (name, mode)
        da, db = _aligned_patristic(ours, ref["newick"])
        assert np.abs(da - db).max() < 4e-8 * m.shape[0], (name, mode)       # "%.8f" text on both sides
        if MODES[mode]["nni"]:
            assert st["n_nni"] == ref["n_nni"], (name, mode)
        if MODES[mode]["spr"]:
            assert st["n_spr"] == ref["n_spr"], (name, mode)
        if mode != "none":
            assert abs(st["length_own"] - ref["length_own"]) < 5e-7, name    # FastME's BIONJ sums in another order
            assert abs(st["length_start"] - ref["length_start"]) < 2e-8, name
        moves += st["n_nni"] + st["n_spr"]
    if mode != "none":
        assert moves > 100      # the fixtures do exercise the searches


def test_reference_trees_of_the_test_alignments():
    """north star: 'FastME trees built from them must have identical topology on data/testdata'.  Here with our
    own tree builder on the reference's matrices against the reference's FastME trees: identical after
    collapsing zero-length branches on 20/20, strictly identical on all but the alignments with several
    identical sequences (where equal-length resolutions exist and FastME's own pick depends on fp noise)."""
    from phyloformer_b200.data import load_alignment_idx
    ref = dict(np.load(os.path.join(GOLDEN, "ref_testdata_pf.npz")))
    trees = json.load(open(os.path.join(GOLDEN, "ref_trees_pf.json")))
    strict = 0
    for stem in sorted(ref):
        _, ids = load_alignment_idx(os.path.join(GOLDEN, "msas", stem + ".fa"))
        m = phylip_rounded(_tri(ref[stem].astype(np.float32)))
        ours = bme_tree(m, ids)
        assert rf_distance(ours, trees[stem], min_length=1e-8) == 0, stem
        strict += rf_distance(ours, trees[stem]) != 0
    assert strict <= 1, strict


def _pauplin_length(newick, dm, ids):
    """Pauplin's formula, brute force: sum over leaf pairs of d_ij * 2^(1 - edges on the path i..j)."""
    root = parse_newick(newick)
    paths = {}

    def walk(node, trail):
        here = trail + [id(node)]
        if not node.children:
            paths[node.name] = here
        for c in node.children:
            walk(c, here)

    walk(root, [])
    total = 0.0
    for a, b in itertools.combinations(range(len(ids)), 2):
        pa, pb = paths[ids[a]], paths[ids[b]]
        k = 0
        while k < min(len(pa), len(pb)) and pa[k] == pb[k]:
            k += 1
        edges = (len(pa) - k) + (len(pb) - k)
        total += dm[a, b] * 2.0 ** (1 - edges)
    return total


def test_balanced_lengths_sum_to_pauplins_formula(fastme_cases):
    """The balanced branch lengths of any topology add up to Pauplin's tree length, and the search never ends on a
    longer tree than it started from."""
    mats, _ = fastme_cases
    for name in ("rand8_0.3_0", "rand12_0.2_1", "rand25_0.1_0", "0_20_tips_noise0.2"):
        m = mats[name]
        ids = [f"T{i + 1}" for i in range(m.shape[0])]
        start, st0 = bme_tree(m, ids, nni=False, spr=False, return_stats=True)
        assert abs(_pauplin_length(start, m, ids) - st0["length_start"]) < 1e-9
        for kw, key in ((dict(nni=True, spr=False), "length_nni"), (dict(nni=False, spr=True), "length_spr")):
            tree, st = bme_tree(m, ids, return_stats=True, **kw)
            assert st[key] <= st["length_start"] + 1e-12
            if st["kept"] != "start":       # a search result carries balanced branch lengths
                assert abs(_pauplin_length(tree, m, ids) - st[key]) < 1e-9
                lens = [float(x) for x in re.findall(r":(-?[0-9.]+)", tree)]
                assert abs(sum(lens) - st[key]) < 1e-6            # printed with 8 digits


def test_additive_matrices_are_recovered():
    """On an additive (tree) metric BIONJ already returns the tree: no move improves it, the branch lengths are
    the tree's own."""
    trees = json.load(open(os.path.join(GOLDEN, "ref_trees_pf.json")))
    for stem in ("0_20_tips", "2_30_tips"):
        names, dm = patristic_distances(trees[stem])
        ours, st = bme_tree(np.asarray(dm), names, return_stats=True)
        assert rf_distance(ours, trees[stem], min_length=1e-9) == 0
        assert st["n_nni"] == 0 and st["n_spr"] == 0
        da, db = _aligned_patristic(ours, trees[stem])
        assert np.abs(da - db).max() < 1e-6


def test_small_and_degenerate_inputs():
    from phyloformer_b200 import _cabi
    assert bme_tree(np.zeros((1, 1)), ["a"]) == "a;"
    assert bme_tree(np.array([[0, 2.0], [2.0, 0]]), ["a", "b c"]) == "(a:1.00000000,'b c':1.00000000);"
    three = np.array([[0, 3.0, 4.0], [3.0, 0, 5.0], [4.0, 5.0, 0]])
    assert bme_tree(three, ["x", "y", "z"]) == "(x:1.00000000,y:2.00000000,z:3.00000000);"
    four = np.array([[0, 2, 4, 4.0], [2, 0, 4, 4], [4, 4, 0, 2], [4, 4, 2, 0]])
    t = bme_tree(four, list("abcd"))
    assert bipartitions(t)[0] == {frozenset("cd")} or bipartitions(t)[0] == {frozenset("ab")}
    zeros = bme_tree(np.zeros((6, 6)), list("abcdef"))      # all sequences identical: any topology, zero lengths
    assert len(bipartitions(zeros)[1]) == 6 and set(re.findall(r":(-?[0-9.]+)", zeros)) == {"0.00000000"}
    with pytest.raises(ValueError):
        bme_tree(np.zeros((3, 3)), ["a", "b"])
    with pytest.raises(_cabi.PfError):
        bme_tree(np.full((3, 3), np.nan), ["a", "b", "c"])
    with pytest.raises(_cabi.PfError, match="exceed the limit"):      # PF_BME_MAX_TAXA: the n^2 table is refused, not attempted
        bme_tree(np.zeros((1001, 1001)), [str(i) for i in range(1001)])


def test_nj_start_tree_option():
    rng = np.random.default_rng(3)
    x = rng.random((15, 4))
    dm = np.abs(x[:, None, :] - x[None, :, :]).sum(-1)
    ids = [f"t{i}" for i in range(15)]
    a = bme_tree(dm, ids, nj_start=True)
    b = bme_tree(dm, ids)
    assert len(bipartitions(a)[1]) == 15 and len(bipartitions(b)[1]) == 15
