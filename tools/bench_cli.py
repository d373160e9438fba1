#!/usr/bin/env python
"""End-to-end CLI throughput (FASTA directory -> PHYLIP files) on the GPU box.

    python tools/bench_cli.py [--n 100] [--L 500] [--files 64] [--trees]

Writes `files` synthetic FASTA alignments of one shape into a temp directory, runs
`infer_alns.main` on it twice (first run pays library load and lazy initialisation) and prints
one JSON line: MSAs/s of the second run, wall clock, with everything inside (parse, H2D,
forward, symmetrise, D2H, '%.10f' formatting, file writes).
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import infer_alns  # noqa: E402

ALPHABET = "ARNDCQEGHILKMFPSTWYVX-"


def write_fasta(path, n, L, rng):
    codes = rng.integers(0, 20, size=(n, L))
    lut = np.frombuffer(ALPHABET.encode(), dtype=np.uint8)
    with open(path, "wb") as f:
        for i in range(n):
            f.write(b">seq%d\n" % i)
            f.write(lut[codes[i]].tobytes() + b"\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--weights", default=os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "ckpt_pf.pt"))
    ap.add_argument("--n", type=int, default=100)
    ap.add_argument("--L", type=int, default=500)
    ap.add_argument("--files", type=int, default=64)
    ap.add_argument("--trees", action="store_true")
    ap.add_argument("--bme", action="store_true", help="also build the BME (FastME-equivalent) trees: infer_alns.py -b")
    a = ap.parse_args()
    rng = np.random.default_rng(7)
    with tempfile.TemporaryDirectory() as tmp:
        src, dst = os.path.join(tmp, "msas"), os.path.join(tmp, "out")
        os.makedirs(src)
        for k in range(a.files):
            write_fasta(os.path.join(src, f"aln{k:05d}.fa"), a.n, a.L, rng)
        argv = [a.weights, src, "-o", dst] + (["-t"] if a.trees else []) + (["-b"] if a.bme else [])
        times = []
        for _ in range(2):
            t0 = time.perf_counter()
            infer_alns.main(argv)
            times.append(time.perf_counter() - t0)
        n_out = len([f for f in os.listdir(dst) if f.endswith(".phy")])
    print(json.dumps({"metric": "cli_msas_per_s", "value": a.files / times[1], "n": a.n, "L": a.L, "files": a.files,
                      "trees": a.trees, "bme_trees": a.bme, "first_run_s": times[0], "second_run_s": times[1], "phy_written": n_out}))


if __name__ == "__main__":
    main()
