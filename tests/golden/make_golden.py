#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference, CPU is enough, ~2 min):

    python tests/golden/make_golden.py

The reference ships no tests / known answers for the hot path, so
every vector here is an output of the reference's own `phyloformer.model.Phyloformer`
(fp32, CPU, torch of this image) on fixed inputs.  The files written:

  ckpt_pf.pt, ckpt_pf_indel.pt   re-serialised checkpoints in the reference's on-disk
                                 layout ({'state_dict': {'model.*'}, 'hyper_parameters'}),
                                 including a (small) stale 'model.seq2pair' entry
  msas/*.fa                      the 20 test alignments (data/testdata/msas)
  ref_testdata_pf.npz            reference distances for those 20 MSAs with pf.ckpt
  ref_phylip_*.phy               reference PHYLIP text for two of them (infer_alns.py:14-25)
  ref_trees_pf.json              FastME (--nni --spr) Newick trees built from the reference
                                 PHYLIP files (README.md:85-92)
  ref_small_taps.npz             hooked intermediates, n=5 L=12
  ref_cases.npz                  small end-to-end cases: soft (non one-hot) input with B=2,
                                 gapped MSA with pf_indel, n=2 and L=1 edge shapes, B=3 batch
  ref_load_alignment.npz         the reference load_alignment() tensor for one file
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.modules.setdefault("dendropy", types.ModuleType("dendropy"))  # data.py:3, unused on this path
sys.path.insert(0, REF)

from phyloformer.model import Phyloformer  # noqa: E402  (the reference's)
from phyloformer.data import load_alignment  # noqa: E402
import importlib.util  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_infer", os.path.join(REF, "infer_alns.py"))
ref_infer = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_infer)

spec = importlib.util.spec_from_file_location(
    "pf_oracle", os.path.join(HERE, "..", "..", "oracle", "pf_oracle.py"))
pf_oracle = importlib.util.module_from_spec(spec)   # by path: the repo root also holds a
spec.loader.exec_module(pf_oracle)                  # `phyloformer` shim that must not shadow REF's


def load_ref_model(name):
    ckpt = torch.load(os.path.join(REF, "models", name), map_location="cpu")
    params = dict(ckpt["hyper_parameters"])
    params["device"] = "cpu"
    m = Phyloformer(**params)
    m.load_state_dict({k.replace("model.", ""): v for k, v in ckpt["state_dict"].items()
                       if k != "model.seq2pair"}, strict=False)
    return m.eval(), ckpt


def main():
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    model, ckpt = load_ref_model("pf.ckpt")
    model_indel, ckpt_indel = load_ref_model("pf_indel.ckpt")

    # --- checkpoints, re-serialised in the reference layout --------------------------
    for name, ck in (("ckpt_pf.pt", ckpt), ("ckpt_pf_indel.pt", ckpt_indel)):
        sd = {k: v.clone() for k, v in ck["state_dict"].items() if k != "model.seq2pair"}
        out = {"state_dict": {"model.seq2pair": torch.zeros(3, 3)}, "hyper_parameters": dict(ck["hyper_parameters"])}
        out["state_dict"].update(sd)
        torch.save(out, os.path.join(HERE, name))

    # --- the 20 test MSAs ------------------------------------------------------------
    msadir = os.path.join(HERE, "msas")
    os.makedirs(msadir, exist_ok=True)
    dists, trees = {}, {}
    fastme = os.path.join(REF, "bin", "bin_linux", "fastme")
    with torch.no_grad(), tempfile.TemporaryDirectory() as tmp:
        for fn in sorted(os.listdir(os.path.join(REF, "data/testdata/msas"))):
            src = os.path.join(REF, "data/testdata/msas", fn)
            shutil.copyfile(src, os.path.join(msadir, fn))
            os.chmod(os.path.join(msadir, fn), 0o644)
            aln, ids = load_alignment(src)
            d = model(aln[None, :].float())
            stem = fn[:-3]
            dists[stem] = d.numpy().astype(np.float32)
            _, phy = ref_infer.vec_to_phylip(d, ids)
            if stem in ("0_20_tips", "3_50_tips"):
                with open(os.path.join(HERE, f"ref_phylip_{stem}.phy"), "w") as fh:
                    fh.write(phy)
            if stem == "0_20_tips":
                np.savez_compressed(os.path.join(HERE, "ref_load_alignment.npz"),
                                    aln=aln.numpy().astype(np.int8), ids=np.array(ids))
            p = os.path.join(tmp, stem + ".phy")
            with open(p, "w") as fh:
                fh.write(phy)
            subprocess.run([fastme, "-i", p, "-o", p + ".nwk", "--nni", "--spr"], check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=tmp)
            trees[stem] = open(p + ".nwk").read().strip()
            print(stem, d.shape, float(d.min()), float(d.max()), flush=True)
    np.savez_compressed(os.path.join(HERE, "ref_testdata_pf.npz"), **dists)
    with open(os.path.join(HERE, "ref_trees_pf.json"), "w") as fh:
        json.dump(trees, fh, indent=0)

    # --- hooked intermediates at a tiny shape ---------------------------------------
    taps = {}
    idx = pf_oracle.synth_msa(5, 12, seed=11)
    x = pf_oracle.msa_to_onehot(idx)
    hooks = []
    hooks.append(model.attention_blocks[0].register_forward_pre_hook(
        lambda m, a: taps.__setitem__("x0", a[0].permute(0, 2, 3, 1).contiguous().numpy())))
    for b, blk in enumerate(model.attention_blocks):
        hooks.append(blk.row_attention.register_forward_hook(  # out (B,P,L,64)
            lambda m, a, o, b=b: taps.__setitem__(f"b{b}.row_attn", o.contiguous().numpy())))
        hooks.append(blk.col_attention.register_forward_hook(  # out (B,L,P,64)
            lambda m, a, o, b=b: taps.__setitem__(f"b{b}.col_attn", o.transpose(1, 2).contiguous().numpy())))
        hooks.append(blk.ffn.register_forward_hook(            # out (B,64,P,L)
            lambda m, a, o, b=b: taps.__setitem__(f"b{b}.ffn_out", o.permute(0, 2, 3, 1).contiguous().numpy())))
        hooks.append(blk.register_forward_hook(
            lambda m, a, o, b=b: taps.__setitem__(f"b{b}.out", o.permute(0, 2, 3, 1).contiguous().numpy())))
    with torch.no_grad():
        d = model(x)
    for h in hooks:
        h.remove()
    taps["idx"] = idx.numpy()
    taps["dist"] = d.numpy()
    np.savez_compressed(os.path.join(HERE, "ref_small_taps.npz"), **taps)

    # --- small end-to-end cases -------------------------------------------------------
    cases = {}
    with torch.no_grad():
        g = torch.Generator().manual_seed(5)
        xs = torch.rand((2, 22, 16, 5), generator=g) * 2 - 0.5     # soft, B=2
        cases["soft_x"] = xs.numpy()
        cases["soft_d"] = model(xs).numpy()
        gi = pf_oracle.synth_msa(8, 40, seed=21, kind="gapped")
        cases["gap_idx"] = gi.numpy()
        cases["gap_d_indel"] = model_indel(pf_oracle.msa_to_onehot(gi)).numpy()
        cases["gap_d_pf"] = model(pf_oracle.msa_to_onehot(gi)).numpy()
        e1 = pf_oracle.synth_msa(2, 7, seed=31)                    # P == 1: squeeze -> 0-dim
        cases["n2_idx"] = e1.numpy()
        cases["n2_d"] = model(pf_oracle.msa_to_onehot(e1)).numpy()
        e2 = pf_oracle.synth_msa(3, 1, seed=32)                    # L == 1
        cases["l1_idx"] = e2.numpy()
        cases["l1_d"] = model(pf_oracle.msa_to_onehot(e2)).numpy()
        e3 = pf_oracle.synth_msa(7, 33, seed=33, B=3)              # batch
        cases["b3_idx"] = e3.numpy()
        cases["b3_d"] = model(pf_oracle.msa_to_onehot(e3)).numpy()
        e4 = pf_oracle.synth_msa(12, 130, seed=34, kind="uniform")  # ragged vs tile sizes
        cases["u12_idx"] = e4.numpy()
        cases["u12_d"] = model(pf_oracle.msa_to_onehot(e4)).numpy()
        # duplicate sequences: rows 0, 3 and 5 identical
        e5 = pf_oracle.synth_msa(9, 50, seed=35)
        e5[0, 3] = e5[0, 0]
        e5[0, 5] = e5[0, 0]
        cases["dup_idx"] = e5.numpy()
        cases["dup_d"] = model(pf_oracle.msa_to_onehot(e5)).numpy()
    np.savez_compressed(os.path.join(HERE, "ref_cases.npz"), **cases)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
