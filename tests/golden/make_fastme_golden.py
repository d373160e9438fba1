#!/usr/bin/env python
"""Known answers for the host-side BME tree search (pf_bme_tree), written by the FastME binary itself.

Runs in the build container only (needs the reference's bin/bin_linux/fastme, staged by tools/stage_ref.sh under
the git-ignored baseline/_ref/bin/): for a set of distance matrices -- the 20 reference matrices with
multiplicative noise (so that the BIONJ start tree is not yet optimal and NNIs / SPRs happen) and random
tree-like matrices -- it stores the matrix exactly as FastME parsed it ('%.10f' text), FastME's Newick
output for `-m I` alone, `--nni`, `--spr` and `--nni --spr`, and the move counts and tree lengths FastME
reports.  Output: tests/golden/fastme_cases.npz (matrices) + tests/golden/fastme_cases.json (trees, stats).

    python tests/golden/make_fastme_golden.py
"""
import json
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
FASTME = os.path.join(ROOT, "baseline", "_ref", "bin", "fastme")
MODES = {"none": [], "nni": ["--nni"], "spr": ["--spr"], "both": ["--nni", "--spr"]}


def tri(v):
    n = int(round((1 + (1 + 8 * len(v)) ** 0.5) / 2))
    m = np.zeros((n, n))
    m[np.triu_indices(n, 1)] = v
    return m + m.T


def random_tree_matrix(n, rng, noise):
    """Path-length matrix of a random tree grown by attaching leaves to random edges, times (1 + noise)."""
    edges = [(0, n, rng.uniform(.01, .3)), (1, n, rng.uniform(.01, .3)), (2, n, rng.uniform(.01, .3))]
    nxt = n + 1
    for leaf in range(3, n):
        k = int(rng.integers(len(edges)))
        a, b, l = edges[k]
        f = rng.uniform(0.2, 0.8)
        edges[k] = (a, nxt, l * f)
        edges.append((nxt, b, l * (1 - f)))
        edges.append((leaf, nxt, rng.uniform(.01, .3)))
        nxt += 1
    import scipy.sparse as sp
    import scipy.sparse.csgraph as cg
    g = sp.lil_matrix((nxt, nxt))
    for a, b, l in edges:
        g[a, b] = g[b, a] = l
    d = cg.shortest_path(g.tocsr(), directed=False)[:n, :n]
    e = np.triu(rng.normal(0, noise, size=(n, n)), 1)
    return np.abs(d * (1 + e + e.T))


def fastme(text, flags, tmp):
    p = os.path.join(tmp, "m.phy")
    open(p, "w").write(text)
    subprocess.run([FASTME, "-i", p, "-o", p + ".nwk"] + flags, check=True, cwd=tmp,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    stat = open(p + "_fastme_stat.txt").read()
    nni = re.search(r"Performed (\d+) NNI", stat)
    spr = re.search(r"Performed (\d+) SPR", stat)
    own = re.search(r"Tree length is ([0-9.]+)", stat)
    start = re.search(r"Before (?:NNI|SPR):\s+tree length is ([0-9.]+)", stat)
    return {"newick": open(p + ".nwk").read().strip(), "n_nni": int(nni.group(1)) if nni else 0,
            "n_spr": int(spr.group(1)) if spr else 0, "length_own": float(own.group(1)) if own else None,
            "length_start": float(start.group(1)) if start else None}


def main():
    if not os.path.exists(FASTME):
        sys.exit("fastme binary not staged (run tools/stage_ref.sh in the build container)")
    rng = np.random.default_rng(20260417)
    ref = dict(np.load(os.path.join(HERE, "ref_testdata_pf.npz")))
    mats = {}
    for k in sorted(ref):
        m = tri(ref[k].astype(np.float64))
        n = m.shape[0]
        if n > 40:
            continue
        for s in (0.05, 0.2):
            e = np.triu(rng.normal(0, s, size=(n, n)), 1)
            mats[f"{k}_noise{s}"] = np.abs(m * (1 + e + e.T))
    for n, noise, reps in ((4, .3, 2), (5, .3, 2), (8, .3, 3), (12, .2, 3), (25, .1, 2), (60, .1, 2)):
        for r in range(reps):
            mats[f"rand{n}_{noise}_{r}"] = random_tree_matrix(n, rng, noise)
    out_m, out_j = {}, {}
    with tempfile.TemporaryDirectory() as tmp:
        for k, m in mats.items():
            n = m.shape[0]
            lines = [f"{n}"] + [f"T{i + 1} " + " ".join(f"{x:.10f}" for x in m[i]) for i in range(n)]
            parsed = np.array([[float(f"{x:.10f}") for x in row] for row in m])
            out_m[k] = parsed
            out_j[k] = {mode: fastme("\n".join(lines) + "\n", flags, tmp) for mode, flags in MODES.items()}
    np.savez_compressed(os.path.join(HERE, "fastme_cases.npz"), **out_m)
    json.dump(out_j, open(os.path.join(HERE, "fastme_cases.json"), "w"), indent=0, sort_keys=True)
    moves = sum(v["both"]["n_nni"] + v["both"]["n_spr"] for v in out_j.values())
    print(f"{len(out_m)} matrices, {moves} NNIs + SPRs in the `--nni --spr` runs")


if __name__ == "__main__":
    main()
