"""Minimal Newick reader and bipartition (Robinson-Foulds) comparison for the topology gate.

Only what the FastME output needs: nested parentheses, leaf names, optional branch lengths and
internal support labels.  Trees are treated as unrooted."""
from typing import FrozenSet, List, Optional, Set, Tuple


class _Node:
    __slots__ = ("children", "name", "length")

    def __init__(self):
        self.children: List["_Node"] = []
        self.name: Optional[str] = None
        self.length: Optional[float] = None


def parse_newick(text: str) -> _Node:
    s = text.strip()
    if s.endswith(";"):
        s = s[:-1]
    pos = 0

    def parse() -> _Node:
        nonlocal pos
        node = _Node()
        if s[pos] == "(":
            pos += 1
            while True:
                node.children.append(parse())
                if s[pos] == ",":
                    pos += 1
                    continue
                if s[pos] == ")":
                    pos += 1
                    break
                raise ValueError(f"bad newick at {pos}: {s[pos:pos + 20]!r}")
        while pos < len(s) and s[pos] in " \t\n":
            pos += 1
        if pos < len(s) and s[pos] == "'":           # quoted label: '' is an embedded quote
            pos += 1
            chars = []
            while True:
                if pos >= len(s):
                    raise ValueError("unterminated quoted label")
                if s[pos] == "'":
                    if pos + 1 < len(s) and s[pos + 1] == "'":
                        chars.append("'")
                        pos += 2
                        continue
                    pos += 1
                    break
                chars.append(s[pos])
                pos += 1
            label = "".join(chars)
        else:
            start = pos
            while pos < len(s) and s[pos] not in ",():;":
                pos += 1
            label = s[start:pos].strip()
        if label and not node.children:
            node.name = label
        if pos < len(s) and s[pos] == ":":
            pos += 1
            start = pos
            while pos < len(s) and s[pos] not in ",()":
                pos += 1
            node.length = float(s[start:pos])
        return node

    return parse()


def bipartitions(text: str, min_length: Optional[float] = None) -> Tuple[Set[FrozenSet[str]], Set[str]]:
    """Non-trivial splits of the unrooted tree, each normalised to the side not containing the
    alphabetically first leaf.  Internal branches with length <= min_length are collapsed."""
    root = parse_newick(text)
    leaves: Set[str] = set()
    splits: Set[FrozenSet[str]] = set()

    def walk(node: _Node) -> FrozenSet[str]:
        if not node.children:
            leaves.add(node.name)
            return frozenset([node.name])
        below = frozenset().union(*[walk(c) for c in node.children])
        if node is not root:
            if min_length is None or node.length is None or node.length > min_length:
                splits.add(below)
        return below

    walk(root)
    ref = min(leaves)
    out = set()
    for sp in splits:
        side = sp if ref not in sp else frozenset(leaves - sp)
        if 1 < len(side) < len(leaves) - 1:
            out.add(side)
    return out, leaves


def rf_distance(a: str, b: str, min_length: Optional[float] = None) -> int:
    """Robinson-Foulds distance (number of splits in exactly one tree)."""
    sa, la = bipartitions(a, min_length)
    sb, lb = bipartitions(b, min_length)
    if la != lb:
        raise ValueError("trees have different leaf sets")
    return len(sa ^ sb)


def patristic_distances(text: str):
    """(leaf names, matrix of path lengths between leaves) of a Newick tree with branch lengths."""
    import numpy as np
    root = parse_newick(text)
    names, paths = [], []  # per leaf: list of (node id, length) up to the root

    def walk(node, trail):
        here = trail + [(id(node), node.length or 0.0)]
        if not node.children:
            names.append(node.name)
            paths.append(here)
        for c in node.children:
            walk(c, here)

    walk(root, [])
    n = len(names)
    dm = np.zeros((n, n))
    for a in range(n):
        da = {k: i for i, (k, _) in enumerate(paths[a])}
        for b in range(a + 1, n):
            # deepest common node
            common = max(i for i, (k, _) in enumerate(paths[b]) if k in da and da[k] == i and paths[a][i][0] == k)
            la = sum(l for _, l in paths[a][common + 1:])
            lb = sum(l for _, l in paths[b][common + 1:])
            dm[a, b] = dm[b, a] = la + lb
    return names, dm
