"""FASTA -> tensors for the inference path (mirrors reference phyloformer/data.py:7-31).

`load_alignment` keeps the reference's return value (int64 one-hot of shape (22, L, n) and
the taxon ids).  `load_alignment_idx` is the compact form the CUDA path consumes directly
((n, L) uint8 residue codes) without the 22x one-hot expansion.  dendropy is not imported:
it is only needed for training labels, which are outside this path.
"""
import numpy as np
import torch

ALPHABET = b"ARNDCQEGHILKMFPSTWYVX-"
LOOKUP = {char: index for index, char in enumerate(ALPHABET)}
_LUT = np.full(256, 255, dtype=np.uint8)
for _i, _c in enumerate(ALPHABET):
    _LUT[_c] = _i


def load_alignment_idx(filepath):
    """Returns ((n, L) uint8 tensor of residue codes, ids). KeyError on a residue outside
    ALPHABET, like the reference (data.py:26)."""
    ids, seqs, cur = [], [], None
    with open(filepath, "rb") as aln:
        for line in aln:
            line = line.strip()
            if line.startswith(b">"):
                ids.append(line[1:].decode("utf8"))
                cur = []
                seqs.append(cur)
            elif line:
                cur.append(line)
    rows = []
    for parts in seqs:
        raw = np.frombuffer(b"".join(parts), dtype=np.uint8)
        codes = _LUT[raw]
        bad = np.nonzero(codes == 255)[0]
        if bad.size:
            raise KeyError(chr(int(raw[bad[0]])))   # same key the reference's LOOKUP[char] raises
        rows.append(codes)
    if len({len(r) for r in rows}) > 1:
        raise ValueError(f"{filepath}: sequences have different lengths")
    return torch.from_numpy(np.stack(rows)) if rows else torch.zeros((0, 0), dtype=torch.uint8), ids


def load_alignment(filepath):
    """Reference-compatible: (int64 one-hot (22, L, n), ids)."""
    idx, ids = load_alignment_idx(filepath)
    seqs = torch.nn.functional.one_hot(idx.long(), num_classes=len(ALPHABET)).permute(2, 1, 0)
    return seqs, ids
