#!/usr/bin/env python
"""Infer evolutionary distances with Phyloformer on a B200.

Same command line as the reference's infer_alns.py (infer_alns.py:42-60):

    python infer_alns.py WEIGHTS ALNDIR -o OUTDIR [-t]

For every .fa/.fasta file in ALNDIR writes OUTDIR/<stem>.phy (PHYLIP distance matrix,
'%.10f', same text as the reference's vec_to_phylip) and, with -t, a neighbour-joining tree
(needs scikit-bio, like the reference).  Differences under the hood: the alignment goes to
the GPU as (n, L) uint8 residue codes instead of a 22-channel fp32 one-hot, alignments of the
same shape are run as one batched forward (the reference loops file by file,
infer_alns.py:97-112), the symmetric matrix is assembled on the device, the text is formatted
in one vectorised pass, and -t falls back to a built-in neighbour joining when scikit-bio is
not installed.  Extra environment knobs: PF_PRECISION=fp32|bf16x3|bf16,
PF_MAX_BATCH_TOKENS (pair*sites per batched call, default 2e7).
"""
import argparse
import os
from glob import glob
from pathlib import Path

import numpy as np
import torch
from tqdm import tqdm

from phyloformer.data import load_alignment_idx
from phyloformer.model import Phyloformer


def matrix_to_phylip(dm: np.ndarray, ids) -> str:
    """(n,n) matrix -> PHYLIP text, byte-identical to the reference (infer_alns.py:19-23)."""
    n = len(ids)
    lines = [f"{n}"]
    dm64 = dm.astype(np.float64)
    for name, row in zip(ids, dm64):
        lines.append(f"{name} " + " ".join(["%.10f" % v for v in row]))
    return "\n".join(lines) + "\n"


def vec_to_phylip(preds: torch.Tensor, ids, model=None):
    """Reference-compatible helper: returns (dm, text)."""
    n = len(ids)
    if model is not None and preds.is_cuda:
        dm = model.distance_matrix(preds, n)[0]
    else:
        dm = torch.zeros((n, n)).type_as(preds)
        i = torch.triu_indices(row=n, col=n, offset=1)
        dm[i[0], i[1]] = preds
        dm = dm + dm.T
    return dm, matrix_to_phylip(dm.detach().cpu().numpy(), ids)


def has_fasta_ext(alnpath):
    return alnpath.lower().endswith(".fa") or alnpath.lower().endswith(".fasta")


def load_model(weights, device):
    ckpt = torch.load(weights, map_location=device)
    params = dict(ckpt["hyper_parameters"])
    params["device"] = device
    model = Phyloformer(**params)
    model.load_state_dict(
        {k.replace("model.", ""): v for k, v in ckpt["state_dict"].items() if k != "model.seq2pair"},
        strict=False,
    )
    return model.to(device).eval()


def main(argv=None):
    parser = argparse.ArgumentParser(description="Infer evolutionnary distances with PhyloFormer")
    parser.add_argument("weights", help="Path to model weights to use")
    parser.add_argument("alndir", help="Path to directory containing alignments to infer")
    parser.add_argument("--outdir", "-o", default=None, required=False,
                        help="Path to directory where inferred distance matrices will be written")
    parser.add_argument("--trees", "-t", action="store_true", help="Output NJ trees as well as matrices")
    args = parser.parse_args(argv)

    nj = None
    if args.trees:
        try:
            from skbio import DistanceMatrix
            from skbio.tree import nj as sk_nj
            nj = lambda dm, ids: str(sk_nj(DistanceMatrix(dm, ids=ids)))  # noqa: E731
        except ImportError:
            from phyloformer_b200.nj import neighbor_joining
            nj = lambda dm, ids: neighbor_joining(dm, ids) + "\n"          # noqa: E731
    if not torch.cuda.is_available():
        raise RuntimeError("infer_alns.py (B200 build) needs a CUDA device; there is no CPU fallback")
    if args.outdir is None:
        parser.error("--outdir/-o is required")
    device = "cuda"
    model = load_model(args.weights, device)

    in_dir = os.path.abspath(args.alndir)
    out_dir = os.path.abspath(args.outdir)
    os.makedirs(out_dir, exist_ok=True)

    # parse everything first (host), bucket by alignment shape
    paths = sorted(glob(f"{in_dir}/*"))
    for alnpath in paths:
        if not has_fasta_ext(alnpath):
            raise ValueError("Input files must be fasta files (.fa or .fasta). Got " f"{alnpath}")
    buckets = {}
    for alnpath in paths:
        idx, ids = load_alignment_idx(alnpath)
        buckets.setdefault(tuple(idx.shape), []).append((alnpath, idx, ids))
    max_tokens = float(os.environ.get("PF_MAX_BATCH_TOKENS", 2e7))

    with torch.no_grad(), tqdm(total=len(paths)) as bar:
        for (n, L), items in buckets.items():
            per_msa = max(1, n * (n - 1) // 2 * L)
            step = max(1, int(max_tokens // per_msa))
            for lo in range(0, len(items), step):
                chunk = items[lo:lo + step]
                batch = torch.stack([it[1] for it in chunk]).to(device, non_blocking=True)   # (B,n,L) uint8
                preds = model.forward_idx(batch, squeeze=False)                              # (B,P)
                mats = model.distance_matrix(preds, n).cpu().numpy()                         # (B,n,n)
                for (alnpath, _, ids), dm in zip(chunk, mats):
                    stem = Path(alnpath).stem
                    with open(os.path.join(out_dir, f"{stem}.phy"), "w") as outfile:
                        outfile.write(matrix_to_phylip(dm, ids))
                    if nj is not None:
                        with open(os.path.join(out_dir, f"{stem}.nj.nwk"), "w") as outfile:
                            outfile.write(nj(dm.astype(np.float64), ids))
                    bar.update(1)


if __name__ == "__main__":
    main()
