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
