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
    """Returns ((n, L) uint8 tensor of residue codes, ids).  Same errors as the reference:
    KeyError(byte) on a residue outside ALPHABET (data.py:26), ValueError for sequences of
    different lengths (the reference's torch.tensor() on ragged lists), IndexError for residues
    before the first header.  The parse itself is the library's host-side pf_parse_fasta (one C
    pass over the file, no GIL, so the CLI's parse pool really runs in parallel)."""
    import ctypes
    from . import _cabi
    with open(filepath, "rb") as aln:
        text = aln.read()
    lib = _cabi.load()
    max_names = text.count(b">") + 1
    codes = np.empty(max(len(text), 1), dtype=np.uint8)
    off = np.empty(max_names, dtype=np.int64)
    ln = np.empty(max_names, dtype=np.int32)
    L, bad = ctypes.c_int32(0), ctypes.c_int32(0)
    n = lib.pf_parse_fasta(text, len(text), codes.ctypes.data, codes.size, ctypes.byref(L), off.ctypes.data,
                           ln.ctypes.data, max_names, ctypes.byref(bad))
    if n == _cabi.PF_ERR_FASTA_RESIDUE:
        raise KeyError(bad.value)
    if n == _cabi.PF_ERR_FASTA_RAGGED:
        raise ValueError(f"{filepath}: sequences have different lengths")
    if n == _cabi.PF_ERR_FASTA_NOHEADER:
        raise IndexError(f"{filepath}: sequence data before the first '>' header")
    if n < 0:
        raise _cabi.PfError(lib.pf_last_error().decode())
    ids = [text[o:o + k].decode("utf8") for o, k in zip(off[:n].tolist(), ln[:n].tolist())]
    if n == 0:
        return torch.zeros((0, 0), dtype=torch.uint8), ids
    return torch.from_numpy(codes[: n * L.value].reshape(n, L.value).copy()), ids


def _load_alignment_idx_py(filepath):
    """Pure-Python restatement of the same parse (tests cross-check the C parser against it)."""
    ids, seqs, cur = [], [], None
    with open(filepath, "rb") as aln:
        for line in aln:
            line = line.strip()
            if line.startswith(b">"):
                ids.append(line[1:].decode("utf8"))
                cur = []
                seqs.append(cur)
            elif line:
                if cur is None:
                    raise IndexError(f"{filepath}: sequence data before the first '>' header")
                cur.append(line)
    rows = []
    for parts in seqs:
        raw = np.frombuffer(b"".join(parts), dtype=np.uint8)
        codes = _LUT[raw]
        bad = np.nonzero(codes == 255)[0]
        if bad.size:
            raise KeyError(int(raw[bad[0]]))   # the reference's LOOKUP has integer (byte) keys
        rows.append(codes)
    if len({len(r) for r in rows}) > 1:
        raise ValueError(f"{filepath}: sequences have different lengths")
    return torch.from_numpy(np.stack(rows)) if rows else torch.zeros((0, 0), dtype=torch.uint8), ids


def load_alignment(filepath):
    """Reference-compatible: (int64 one-hot (22, L, n), ids)."""
    idx, ids = load_alignment_idx(filepath)
    seqs = torch.nn.functional.one_hot(idx.long(), num_classes=len(ALPHABET)).permute(2, 1, 0)
    return seqs, ids
