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
not installed; parsing, device work and file output overlap (run_pipeline).  Extra environment
knobs: PF_PRECISION=fp32|bf16x3|fp16|bf16, PF_MAX_BATCH_TOKENS (pair*sites per batched call,
default 2e7).
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
    """(n,n) matrix -> PHYLIP text, byte-identical to the reference (infer_alns.py:19-23).
    fp32 matrices go through the library's host formatter (pf_format_phylip: exact '%.10f' in
    C, runs without the GIL so the CLI's writer threads overlap); anything else, or names the C
    string interface cannot carry, takes the pure-Python path."""
    n = len(ids)
    if dm.dtype == np.float32 and dm.shape == (n, n) and all("\0" not in str(i) for i in ids):
        import ctypes
        from phyloformer_b200 import _cabi
        lib = _cabi.load()
        dmc = np.ascontiguousarray(dm)
        names = (ctypes.c_char_p * n)(*[str(i).encode("utf8") for i in ids])
        cap = 16 + sum(len(b) + 1 for b in names) + n * n * 14
        while True:
            buf = ctypes.create_string_buffer(cap)
            need = lib.pf_format_phylip(dmc.ctypes.data, n, names, buf, cap)
            if need < 0:
                raise _cabi.PfError(lib.pf_last_error().decode())
            if need <= cap:
                return buf.raw[:need].decode("utf8")
            cap = need
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
    parser.add_argument("--bme-trees", "-b", action="store_true",
                        help="Also output <stem>.bme.nwk: BIONJ + balanced NNI/SPR trees, what the README's "
                             "`fastme -i <stem>.phy --nni --spr` step builds (not a reference option)")
    args = parser.parse_args(argv)

    nj = None
    if args.trees:
        try:
            from skbio import DistanceMatrix
            from skbio.tree import nj as sk_nj
            nj = lambda dm, ids: str(sk_nj(DistanceMatrix(dm, ids=ids)))  # noqa: E731
        except ImportError:
            from phyloformer_b200.nj import neighbor_joining_c
            nj = lambda dm, ids: neighbor_joining_c(dm, ids) + "\n"        # noqa: E731
    bme = None
    if args.bme_trees:
        from phyloformer_b200.bme import bme_tree
        bme = lambda dm, ids: bme_tree(np.round(dm, 10), ids) + "\n"   # noqa: E731  (the '%.10f' values of the .phy text)
    if not torch.cuda.is_available():
        raise RuntimeError("infer_alns.py (B200 build) needs a CUDA device; there is no CPU fallback")
    if args.outdir is None:
        parser.error("--outdir/-o is required")
    device = "cuda"
    model = load_model(args.weights, device)

    in_dir = os.path.abspath(args.alndir)
    out_dir = os.path.abspath(args.outdir)
    os.makedirs(out_dir, exist_ok=True)

    paths = sorted(glob(f"{in_dir}/*"))
    for alnpath in paths:
        if not has_fasta_ext(alnpath):
            raise ValueError("Input files must be fasta files (.fa or .fasta). Got " f"{alnpath}")
    max_tokens = float(os.environ.get("PF_MAX_BATCH_TOKENS", 2e7))
    with torch.no_grad(), tqdm(total=len(paths)) as bar:
        run_pipeline(model, paths, out_dir, nj, max_tokens, bar.update, bme=bme)


def write_outputs(out_dir, chunk, mats, nj, bme=None):
    """Host-side tail for one batched call: PHYLIP text (and optionally the NJ / BME tree) per file."""
    for (alnpath, _, ids), dm in zip(chunk, mats):
        stem = Path(alnpath).stem
        with open(os.path.join(out_dir, f"{stem}.phy"), "w") as outfile:
            outfile.write(matrix_to_phylip(dm, ids))
        if nj is not None:
            with open(os.path.join(out_dir, f"{stem}.nj.nwk"), "w") as outfile:
                outfile.write(nj(dm.astype(np.float64), ids))
        if bme is not None:
            with open(os.path.join(out_dir, f"{stem}.bme.nwk"), "w") as outfile:
                outfile.write(bme(dm.astype(np.float64), ids))
    return len(chunk)


MAX_BATCH = 65535  # alignments per batched call: the batch index is a gridDim.y/z of several kernels


def run_pipeline(model, paths, out_dir, nj, max_tokens, progress=lambda k: None, depth=2, bme=None):
    """Three overlapped stages (the reference does them serially per file, infer_alns.py:97-117):

      parse        FASTA -> (n, L) uint8 codes (C parser), on the submitting thread while the
                   device works on earlier batches
      device       alignments of one shape are stacked into a pinned staging buffer, copied
                   H2D, run as ONE batched forward, symmetrised on the device and copied D2H into
                   a pinned result buffer, all asynchronously on the current stream; up to `depth`
                   batches are in flight before the host waits for the oldest one
      write pool   '%.10f' formatting + file output of finished batches
    """
    from collections import deque
    from concurrent.futures import ThreadPoolExecutor

    device = next(model.parameters()).device
    n_cpu = os.cpu_count() or 1
    inflight = deque()
    writes = []

    def launch(chunk, n, L):
        stage = torch.empty((len(chunk), n, L), dtype=torch.uint8).pin_memory()
        for k, it in enumerate(chunk):
            stage[k].copy_(it[1])
        batch = stage.to(device, non_blocking=True)                      # (B,n,L) uint8
        preds = model.forward_idx(batch, squeeze=False)                  # (B,P)
        mats_dev = model.distance_matrix(preds, n)                       # (B,n,n)
        mats = torch.empty(mats_dev.shape, dtype=mats_dev.dtype).pin_memory()
        mats.copy_(mats_dev, non_blocking=True)
        done = torch.cuda.Event()
        done.record()
        inflight.append((chunk, stage, mats, done))

    def retire(pool):
        chunk, _stage, mats, done = inflight.popleft()
        done.synchronize()
        writes.append(pool.submit(write_outputs, out_dir, chunk, mats.numpy(), nj, bme))

    # Parsing stays on this thread: it is one C pass per file (pf_parse_fasta, ~40 us for 20 x 200)
    # and every device call below is asynchronous, so the GPU works on batch k while this loop
    # parses batch k+1; a parse pool only added GIL hand-offs (measured 0.23 vs 0.04 ms per file).
    # text formatting needs ~4 writer threads; tree building (host-side C, no GIL; ~30 ms per 100-taxon BME tree)
    # gets up to 8 so that the GPU stays the slower side
    with ThreadPoolExecutor(max(1, min(8 if (nj is not None or bme is not None) else 4, n_cpu))) as write_pool:
        buckets = {}
        for alnpath in paths:
            idx, ids = load_alignment_idx(alnpath)
            n, L = idx.shape
            per_msa = max(1, n * (n - 1) // 2 * L)
            step = max(1, min(MAX_BATCH, int(max_tokens // per_msa)))
            items = buckets.setdefault((n, L), [])
            items.append((alnpath, idx, ids))
            if len(items) >= step:
                launch(buckets.pop((n, L)), n, L)
                while len(inflight) > depth:
                    retire(write_pool)
            while writes and writes[0].done():
                progress(writes.pop(0).result())
        for (n, L), items in buckets.items():        # partially filled buckets
            launch(items, n, L)
            while len(inflight) > depth:
                retire(write_pool)
        while inflight:
            retire(write_pool)
        for w in writes:
            progress(w.result())
    model.check_device_error()    # a bounded device-side wait that timed out must not pass silently


if __name__ == "__main__":
    main()
