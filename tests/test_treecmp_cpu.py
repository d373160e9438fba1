"""CPU tests of the Newick/bipartition helper used by the FastME topology gate."""
import json
import os

import numpy as np

from phyloformer_b200.treecmp import bipartitions, rf_distance
from tests._util import GOLDEN


def test_rf_basics():
    a = "((A:1,B:1):1,(C:1,D:1):1,E:1);"
    assert rf_distance(a, "((B:2,A:1):3,E:1,(D:1,C:1):0.5);") == 0          # same unrooted topology
    assert rf_distance(a, "((A:1,C:1):1,(B:1,D:1):1,E:1);") == 4
    assert rf_distance(a, "((A:1,B:1):0,(C:1,D:1):1,E:1);", min_length=1e-8) == 1  # collapsed branch


def test_reference_trees_parse():
    trees = json.load(open(os.path.join(GOLDEN, "ref_trees_pf.json")))
    assert len(trees) == 20
    for stem, nwk in trees.items():
        splits, leaves = bipartitions(nwk)
        n = int(stem.split("_")[1])
        assert len(leaves) == n and len(splits) <= n - 3
        assert rf_distance(nwk, nwk) == 0


def test_neighbor_joining_recovers_additive_trees():
    """NJ is exact on additive (tree) metrics: feed it the path-length matrices of the stored
    reference FastME trees and require RF = 0 (ignoring zero-length branches)."""
    import numpy as np
    from phyloformer_b200.nj import neighbor_joining
    from phyloformer_b200.treecmp import patristic_distances
    trees = json.load(open(os.path.join(GOLDEN, "ref_trees_pf.json")))
    for stem in ("0_20_tips", "2_30_tips", "2_50_tips"):
        names, dm = patristic_distances(trees[stem])
        assert np.allclose(dm, dm.T) and (dm >= 0).all()
        nwk = neighbor_joining(dm, names)
        assert rf_distance(nwk, trees[stem], min_length=1e-9) == 0, stem
    assert neighbor_joining(np.array([[0, 2.0], [2.0, 0]]), ["a", "b"]) == "(a:1.0000000000,b:1.0000000000);"


def test_c_neighbor_joining_matches_python():
    """pf_neighbor_joining (host-only C-ABI entry) against phyloformer_b200/nj.py: same topology and
    branch lengths on additive and noisy matrices, same text in the degenerate cases."""
    import re
    import numpy as np
    from phyloformer_b200.nj import neighbor_joining, neighbor_joining_c
    from phyloformer_b200.treecmp import bipartitions, patristic_distances
    rng = np.random.default_rng(5)
    for n in (3, 4, 7, 20, 61):
        x = rng.random((n, 5))
        dm = np.abs(x[:, None, :] - x[None, :, :]).sum(-1) + rng.random((n, n)) * 0.01
        dm = ((dm + dm.T) / 2).astype(np.float32)
        np.fill_diagonal(dm, 0)
        ids = [f"t{i}" for i in range(n)]
        a, b = neighbor_joining(dm.astype(np.float64), ids), neighbor_joining_c(dm, ids)
        assert bipartitions(a)[0] == bipartitions(b)[0]
        la = [float(v) for v in re.findall(r":([0-9.]+)", a)]
        lb = [float(v) for v in re.findall(r":([0-9.]+)", b)]
        assert len(la) == len(lb) and np.allclose(la, lb, atol=2e-10)
        (na, pa), (nb, pb) = patristic_distances(a), patristic_distances(b)
        assert na == nb and np.abs(pa - pb).max() < 1e-8
    assert neighbor_joining_c(np.array([[0, 2.0], [2.0, 0]]), ["a", "b"]) == "(a:1.0000000000,b:1.0000000000);"
    assert neighbor_joining_c(np.zeros((1, 1)), ["solo"]) == "solo;"


def test_neighbor_joining_known_answer():
    """External known answer: the five-taxon matrix of the neighbour-joining worked example (Saitou & Nei's
    algorithm as in the Wikipedia article; the same matrix is the doctest of `skbio.tree.nj`, the call behind the
    reference's `--trees`, infer_alns.py:120-123), whose published result is
        (d:2, (c:4, (b:3, a:2):3):2, e:1);
    including both Q-matrix ties (a,b)/(d,e) and (u,c)/(d,e), which must resolve to the first pair."""
    from phyloformer_b200.nj import neighbor_joining, neighbor_joining_c
    from phyloformer_b200.treecmp import parse_newick
    dm = np.array([[0, 5, 9, 9, 8], [5, 0, 10, 10, 9], [9, 10, 0, 8, 7], [9, 10, 8, 0, 3], [8, 9, 7, 3, 0]], dtype=np.float64)
    want = "(((a:2.0000000000,b:3.0000000000):3.0000000000,c:4.0000000000):2.0000000000,d:2.0000000000,e:1.0000000000);"
    for fn in (neighbor_joining, neighbor_joining_c):
        nwk = fn(dm, list("abcde"))
        assert nwk == want, (fn.__name__, nwk)
        splits, leaves = bipartitions(nwk)
        assert leaves == set("abcde") and splits == {frozenset("cde"), frozenset("de")}
        # leaf branch lengths of the published tree
        root = parse_newick(nwk)
        got = {}

        def walk(nd):
            if nd.name is not None:
                got[nd.name] = nd.length
            for ch in nd.children:
                walk(ch)
        walk(root)
        assert got == {"a": 2.0, "b": 3.0, "c": 4.0, "d": 2.0, "e": 1.0}
    s1, _ = bipartitions(want)
    s2, _ = bipartitions("(d:2.0,(c:4.0,(b:3.0,a:2.0):3.0):2.0,e:1.0);")        # the published Newick
    assert s1 == s2 and len(s1) == 2


def test_newick_labels_with_reserved_characters_are_quoted():
    """FASTA ids are whole header lines: blanks and Newick punctuation must not break the tree file."""
    from phyloformer_b200.nj import neighbor_joining, neighbor_joining_c, newick_label
    assert newick_label("plain_id.1|x") == "plain_id.1|x"
    assert newick_label("Homo sapiens (human)") == "'Homo sapiens (human)'"
    assert newick_label("it's") == "'it''s'"
    ids = ["sp A", "b:1", "c,d", "(e)", "f'g", "ok"]
    rng = np.random.default_rng(3)
    dm = rng.uniform(0.1, 1.0, (6, 6)); dm = dm + dm.T; np.fill_diagonal(dm, 0)
    for fn in (neighbor_joining, neighbor_joining_c):
        nwk = fn(dm, ids)
        _, leaves = bipartitions(nwk)
        assert leaves == set(ids), (fn.__name__, nwk)
    assert neighbor_joining_c(np.array([[0, 2.0], [2.0, 0]]), ["a b", "c"]) == "('a b':1.0000000000,c:1.0000000000);"
