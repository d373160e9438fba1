"""Pair-axis sharding of one forward across ranks.

Everything on the path is independent per pair except column attention (reference
model.py:97), which sums over pairs at every site.  A rank owns a contiguous range of the
lexicographic pair list; the only exchange is a sum of the (B, L, 72) fp32 column summaries
once per block, plus a gather of the distances at the end.
"""
from typing import List, Tuple


def n_pairs(n: int) -> int:
    return n * (n - 1) // 2


def pair_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced (sizes differ by at most 1) range of pairs for `rank`."""
    P = n_pairs(n)
    base, rem = divmod(P, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_ranges(n: int, world: int) -> List[Tuple[int, int]]:
    return [pair_range(n, r, world) for r in range(world)]


def pair_index(i: int, j: int, n: int) -> int:
    """Index of pair (i<j) in the reference's order (model.py:13-17)."""
    return i * n - i * (i + 1) // 2 + (j - i - 1)


def pair_to_ij(p: int, n: int) -> Tuple[int, int]:
    """Inverse of pair_index (same closed form as the device code)."""
    import math
    b = 2 * n - 1
    i = int((b - math.sqrt(max(b * b - 8 * p, 0))) / 2)
    i = max(i, 0)
    while i > 0 and i * n - i * (i + 1) // 2 > p:
        i -= 1
    while (i + 1) * n - (i + 1) * (i + 2) // 2 <= p:
        i += 1
    return i, i + 1 + (p - (i * n - i * (i + 1) // 2))


def batch_range(B: int, rank: int, world: int) -> Tuple[int, int]:
    """MSA-level sharding for batches of small alignments (no exchange at all)."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
