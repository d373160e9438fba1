#!/usr/bin/env python
"""Per precision mode: parity against the committed reference distances of the 20 test
alignments (max / mean relative error), FastME topology agreement with the reference's trees,
and the forward time at 200 x 1000.  Run on the GPU box:  python tools/prec_report.py
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import infer_alns  # noqa: E402
from phyloformer.data import load_alignment_idx  # noqa: E402
from phyloformer_b200.treecmp import rf_distance  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
FASTME = os.path.join(ROOT, "baseline", "_ref", "bin", "fastme")


def main():
    model = infer_alns.load_model(os.path.join(GOLDEN, "ckpt_pf.pt"), "cuda")
    ref = dict(np.load(os.path.join(GOLDEN, "ref_testdata_pf.npz")))
    ref_trees = json.load(open(os.path.join(GOLDEN, "ref_trees_pf.json")))
    stems = sorted(ref)
    big = torch.randint(0, 20, (1, 200, 1000), dtype=torch.uint8, generator=torch.Generator().manual_seed(1)).cuda()
    for prec in ("fp32", "bf16x3", "fp16", "bf16"):
        model.set_precision(prec)
        mx, mean, strict, collapsed = 0.0, [], 0, 0
        with tempfile.TemporaryDirectory() as tmp, torch.no_grad():
            for stem in stems:
                idx, ids = load_alignment_idx(os.path.join(GOLDEN, "msas", stem + ".fa"))
                d = model.forward_idx(idx.cuda())
                r = np.abs(d.double().cpu().numpy() - ref[stem]) / np.abs(ref[stem])
                mx = max(mx, float(r.max()))
                mean.append(float(r.mean()))
                if os.path.exists(FASTME):
                    _, phy = infer_alns.vec_to_phylip(d, ids, model)
                    p = os.path.join(tmp, stem + ".phy")
                    open(p, "w").write(phy)
                    subprocess.run([FASTME, "-i", p, "-o", p + ".nwk", "--nni", "--spr"], check=True,
                                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=tmp)
                    ours = open(p + ".nwk").read().strip()
                    strict += rf_distance(ours, ref_trees[stem]) != 0
                    collapsed += rf_distance(ours, ref_trees[stem], min_length=1e-8) != 0
            for _ in range(2):
                model.forward_idx(big)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                model.forward_idx(big)
            b.record()
            torch.cuda.synchronize()
        print(json.dumps({"precision": prec, "max_rel_20msas": mx, "mean_rel_20msas": float(np.mean(mean)),
                          "trees_differ_strict": int(strict), "trees_differ_collapsed": int(collapsed),
                          "ms_200x1000": a.elapsed_time(b) / 3}))


if __name__ == "__main__":
    main()
