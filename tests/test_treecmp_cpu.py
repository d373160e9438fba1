"""CPU tests of the Newick/bipartition helper used by the FastME topology gate."""
import json
import os

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
