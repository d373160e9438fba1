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
