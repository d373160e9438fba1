#!/usr/bin/env python
"""pf_bme_tree against the FastME binary on random tree-like matrices (build container only: needs
baseline/_ref/bin/fastme, staged by tools/stage_ref.sh).  Per matrix: both run `--nni --spr`; reports the
Robinson-Foulds distance, the largest difference between the trees' path lengths, move counts and wall time.

    python tools/bme_vs_fastme.py [n ...]          # default sizes: 8 25 60 100 200 500
"""
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_fastme_golden import FASTME, fastme, random_tree_matrix  # noqa: E402
from phyloformer_b200.bme import bme_tree  # noqa: E402
from phyloformer_b200.treecmp import patristic_distances, rf_distance  # noqa: E402


def main():
    if not os.path.exists(FASTME):
        sys.exit("fastme binary not staged (run tools/stage_ref.sh in the build container)")
    sizes = [int(a) for a in sys.argv[1:]] or [8, 25, 60, 100, 200, 500]
    rng = np.random.default_rng(99)
    print("| taxa | noise | RF | max path-length diff | NNIs (ours / FastME) | SPRs (ours / FastME) | kept | ours s | FastME s |")
    print("|---|---|---|---|---|---|---|---|---|")
    with tempfile.TemporaryDirectory() as tmp:
        for n in sizes:
            for noise in (0.05, 0.2):
                m = random_tree_matrix(n, rng, noise)
                text = f"{n}\n" + "".join(f"T{i + 1} " + " ".join(f"{x:.10f}" for x in m[i]) + "\n" for i in range(n))
                parsed = np.array([[float(f"{x:.10f}") for x in row] for row in m])
                ids = [f"T{i + 1}" for i in range(n)]
                t0 = time.perf_counter()
                ref = fastme(text, ["--nni", "--spr"], tmp)
                t1 = time.perf_counter()
                ours, st = bme_tree(parsed, ids, return_stats=True)
                t2 = time.perf_counter()
                na, da = patristic_distances(ours)
                nb, db = patristic_distances(ref["newick"])
                ia, ib = [na.index(x) for x in sorted(na)], [nb.index(x) for x in sorted(nb)]
                diff = np.abs(np.asarray(da)[np.ix_(ia, ia)] - np.asarray(db)[np.ix_(ib, ib)]).max()
                print(f"| {n} | {noise} | {rf_distance(ours, ref['newick'])} | {diff:.1e} | {st['n_nni']} / {ref['n_nni']} | "
                      f"{st['n_spr']} / {ref['n_spr']} | {st['kept']} | {t2 - t1:.2f} | {t1 - t0:.2f} |", flush=True)


if __name__ == "__main__":
    main()
