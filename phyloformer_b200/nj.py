"""Neighbour joining (Saitou & Nei 1987, Studier & Keppler formulation) on a distance matrix.

Replaces the `skbio.tree.nj` call behind `infer_alns.py --trees` (reference infer_alns.py:120-123;
scikit-bio is an optional dependency of the reference and is not installed here).  O(n^3) with
numpy row operations, fine for the alignment sizes the model handles (n up to a few hundred).
Returns an unrooted tree as a Newick string with a trifurcation at the root."""
from typing import List, Sequence

import numpy as np


_NEWICK_RESERVED = set(" \t\r\n,:;()[]'")


def newick_label(name) -> str:
    """A taxon name as a Newick label: FASTA ids are whole header lines (reference data.py:22), so they can hold
    blanks or Newick punctuation; such names are single-quoted with embedded quotes doubled (as the skbio
    writer behind the reference's `--trees` does), all others are written as they are."""
    name = str(name)
    if name == "" or any(c in _NEWICK_RESERVED for c in name):
        return "'" + name.replace("'", "''") + "'"
    return name


def neighbor_joining(dm: np.ndarray, ids: Sequence[str], clip_negative: bool = True) -> str:
    d = np.array(dm, dtype=np.float64)
    ids = [newick_label(i) for i in ids]
    n = d.shape[0]
    if d.shape != (n, n) or n != len(ids):
        raise ValueError("distance matrix and ids do not match")
    if n == 1:
        return f"{ids[0]};"
    if n == 2:
        return f"({ids[0]}:{d[0, 1] / 2:.10f},{ids[1]}:{d[0, 1] / 2:.10f});"
    nodes: List[str] = [str(i) for i in ids]
    active = list(range(n))
    d = d.copy()

    def fmt(x):
        if clip_negative and x < 0:
            x = 0.0
        return f"{x:.10f}"

    while len(active) > 3:
        m = len(active)
        sub = d[np.ix_(active, active)]
        r = sub.sum(axis=1)
        q = (m - 2) * sub - r[:, None] - r[None, :]
        np.fill_diagonal(q, np.inf)
        a, b = np.unravel_index(np.argmin(q), q.shape)
        if a > b:
            a, b = b, a
        ia, ib = active[a], active[b]
        dab = sub[a, b]
        la = 0.5 * dab + (r[a] - r[b]) / (2 * (m - 2))
        lb = dab - la
        new = f"({nodes[ia]}:{fmt(la)},{nodes[ib]}:{fmt(lb)})"
        # distances from the new node to the rest, stored in ia's slot
        dn = 0.5 * (d[ia, :] + d[ib, :] - dab)
        d[ia, :] = dn
        d[:, ia] = dn
        d[ia, ia] = 0.0
        nodes[ia] = new
        active.pop(b)
    i, j, k = active
    li = 0.5 * (d[i, j] + d[i, k] - d[j, k])
    lj = 0.5 * (d[i, j] + d[j, k] - d[i, k])
    lk = 0.5 * (d[i, k] + d[j, k] - d[i, j])
    return f"({nodes[i]}:{fmt(li)},{nodes[j]}:{fmt(lj)},{nodes[k]}:{fmt(lk)});"


def neighbor_joining_c(dm: np.ndarray, ids: Sequence[str]) -> str:
    """The same algorithm through the library's host-side pf_neighbor_joining (C, no GIL): what
    the CLI uses.  `dm` is converted to fp32 (the model's output precision)."""
    import ctypes
    from . import _cabi
    lib = _cabi.load()
    n = len(ids)
    d = np.ascontiguousarray(dm, dtype=np.float32)
    if d.shape != (n, n):
        raise ValueError("distance matrix and ids do not match")
    names = (ctypes.c_char_p * n)(*[str(i).encode("utf8") for i in ids])
    cap = 64 + sum(len(b) for b in names) + 40 * n
    while True:
        buf = ctypes.create_string_buffer(cap)
        need = lib.pf_neighbor_joining(d.ctypes.data, n, names, buf, cap)
        if need < 0:
            raise _cabi.PfError(lib.pf_last_error().decode())
        if need <= cap:
            return buf.raw[:need].decode("utf8")
        cap = need
