"""Trees from distance matrices by balanced minimum evolution: BIONJ start tree, then balanced NNI and
SPR searches -- the tree step of the reference's workflow (README.md:85-92 runs the FastME 2.1.6.4 binary,
`fastme -i x.phy -o x.nwk --nni --spr`, on every matrix this path produces).

The work is done by the library's host-side C++ entry `pf_bme_tree` (csrc/pf_bme.h; no device work, releases
the GIL).  On the 20 reference matrices and on perturbed / random tree-like matrices up to 200 taxa it
reproduces the FastME binary's output: same number of NNIs and SPRs, same topology, same "%.8f" branch
lengths (tools/bme_vs_fastme.py; tests/test_bme_cpu.py holds FastME's own outputs as fixtures)."""
import ctypes
from typing import Sequence

import numpy as np

NNI, SPR, NJ_START = 1, 2, 4


def phylip_rounded(dm) -> np.ndarray:
    """The matrix as a reader of this package's PHYLIP text sees it: every value through '%.10f'
    (infer_alns.py:21-22) and back to double."""
    d = np.asarray(dm, dtype=np.float64)
    return np.array([[float(f"{x:.10f}") for x in row] for row in d], dtype=np.float64)


def bme_tree(dm, ids: Sequence[str], nni: bool = True, spr: bool = True, nj_start: bool = False,
             return_stats: bool = False):
    """Newick text (trifurcating root, '%.8f' branch lengths like FastME's default) for the symmetric
    (n,n) matrix `dm`.  With return_stats also a dict: tree lengths of the start tree (own and balanced
    branch lengths), after the NNI search, after the SPR search, the move counts and which tree was kept."""
    from . import _cabi
    lib = _cabi.load()
    n = len(ids)
    d = np.ascontiguousarray(dm, dtype=np.float64)
    if d.shape != (n, n):
        raise ValueError("distance matrix and ids do not match")
    names = (ctypes.c_char_p * n)(*[str(i).encode("utf8") for i in ids])
    flags = (NNI if nni else 0) | (SPR if spr else 0) | (NJ_START if nj_start else 0)
    stats = np.zeros(7, dtype=np.float64)
    cap = 64 + sum(len(b) + 2 for b in names) + 32 * n
    while True:
        buf = ctypes.create_string_buffer(cap)
        need = lib.pf_bme_tree(d.ctypes.data, n, names, flags, buf, cap, stats.ctypes.data)
        if need < 0:
            raise _cabi.PfError(lib.pf_last_error().decode())
        if need <= cap:
            break
        cap = need
    text = buf.raw[:need].decode("utf8")
    if not return_stats:
        return text
    return text, {"length_own": stats[0], "length_start": stats[1], "length_nni": stats[2], "length_spr": stats[3],
                  "n_nni": int(stats[4]), "n_spr": int(stats[5]), "kept": ("start", "nni", "spr")[int(stats[6])]}
