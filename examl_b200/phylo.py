"""Host-side tree structures for the likelihood path: ExaML's node rings, traversal
descriptors and Newick I/O, restated in Python so that tests and benchmarks can drive the
C-ABI engine on machines where the reference's C host code is not available.

Mirrors (behaviour, not code):
  node / nodeptr rings, `x` orientation flags   axml.h:492-506, getxnode axml.c:461-476
  isTip                                          axml.c:443-459
  hookup                                         axml.c:478-490
  computeTraversalInfo                           newviewGenericSpecial.c:691-811
  traversalInfo                                  axml.h:434-442
"""
from __future__ import annotations

import random
import re
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

TIP_TIP, TIP_INNER, INNER_INNER = 0, 1, 2
DEFAULTZ = 0.9          # axml.h:96
ZMIN = 1.0e-15          # axml.h:110
ZMAX = 1.0 - 1.0e-6     # axml.h:111


class Node:
    """One record of a node ring (a tip has a single record, an inner node three)."""
    __slots__ = ("number", "next", "back", "x", "z")

    def __init__(self, number: int):
        self.number = number
        self.next: Optional["Node"] = None
        self.back: Optional["Node"] = None
        self.x = 0
        self.z: List[float] = []


@dataclass
class Entry:
    """traversalInfo (axml.h:434-442)."""
    tipCase: int
    pNumber: int
    qNumber: int
    rNumber: int
    qz: List[float]
    rz: List[float] = field(default_factory=list)


def hookup(p: Node, q: Node, z: Sequence[float]) -> None:
    p.back, q.back = q, p
    zz = list(z)
    p.z = zz
    q.z = zz          # one shared list per branch: both ends always agree


class Tree:
    """Unrooted binary tree with ExaML numbering: tips 1..mxtips, inner mxtips+1..2*mxtips-2."""

    def __init__(self, mxtips: int, num_branches: int = 1):
        self.mxtips = mxtips
        self.numBranches = num_branches
        self.nodep: List[Optional[Node]] = [None] * (2 * mxtips)
        for i in range(1, mxtips + 1):
            self.nodep[i] = Node(i)
        for i in range(mxtips + 1, 2 * mxtips - 1):
            a, b, c = Node(i), Node(i), Node(i)
            a.next, b.next, c.next = b, c, a
            self.nodep[i] = a
        self.start: Node = self.nodep[1]
        self.names: List[str] = [f"T{i}" for i in range(1, mxtips + 1)]

    # ---- helpers -------------------------------------------------------------------------
    def isTip(self, number: int) -> bool:
        return number <= self.mxtips

    @staticmethod
    def getxnode(p: Node) -> None:
        """Make p the record of its ring that owns the CLV (axml.c:461-476)."""
        p.x = 1
        p.next.x = 0
        p.next.next.x = 0

    def branches(self):
        """Every branch once, as (record, record.back)."""
        seen = set()
        for i in range(1, 2 * self.mxtips - 1):
            n = self.nodep[i]
            ring = [n] if self.isTip(i) else [n, n.next, n.next.next]
            for r in ring:
                if r.back is not None and id(r.z) not in seen:
                    seen.add(id(r.z))
                    yield r, r.back

    def inner_branches_from(self, p: Node):
        """Records p (inner) whose branch p--p.back can serve as a virtual root."""
        return [a for a, b in self.branches()]

    # ---- construction ---------------------------------------------------------------------
    @classmethod
    def random(cls, mxtips: int, seed: int = 20260922, num_branches: int = 1, z: float = DEFAULTZ) -> "Tree":
        """Random topology by sequential random attachment (SURVEY.md 8(d) generator)."""
        rng = random.Random(seed)
        t = cls(mxtips, num_branches)
        zz = [z] * num_branches
        nxt = mxtips + 1
        # three-taxon star
        c = t.nodep[nxt]; nxt += 1
        hookup(c, t.nodep[1], zz); hookup(c.next, t.nodep[2], zz); hookup(c.next.next, t.nodep[3], zz)
        edges = [(c, t.nodep[1]), (c.next, t.nodep[2]), (c.next.next, t.nodep[3])]
        for tip in range(4, mxtips + 1):
            a, b = edges.pop(rng.randrange(len(edges)))
            n = t.nodep[nxt]; nxt += 1
            hookup(n, a, zz); hookup(n.next, b, zz); hookup(n.next.next, t.nodep[tip], zz)
            edges += [(n, a), (n.next, b), (n.next.next, t.nodep[tip])]
        t.start = t.nodep[1]
        return t

    @classmethod
    def from_newick(cls, text: str, names: Sequence[str], num_branches: int = 1,
                    read_branches: bool = False) -> "Tree":
        """Parse an unrooted (trifurcating root) or rooted (bifurcating root) binary Newick tree.
        Like the reference's treeReadLen(..., readBranches=FALSE) (axml.c:2308) branch lengths are
        ignored unless read_branches is set, in which case z = exp(-t)."""
        import math
        name_to_num = {n: i + 1 for i, n in enumerate(names)}
        t = cls(len(names), num_branches)
        t.names = list(names)
        tokens = re.findall(r"[(),;]|:[^(),;]+|[^(),:;\s]+", text.strip())
        pos = 0
        nxt = [t.mxtips + 1]

        def zlist(tok_len: Optional[float]):
            if read_branches and tok_len is not None:
                return [min(max(math.exp(-tok_len), ZMIN), ZMAX)] * num_branches
            return [DEFAULTZ] * num_branches

        def parse_subtree():
            """Returns (dangling record, branch length) for the subtree starting at tokens[pos]."""
            nonlocal pos
            if tokens[pos] == "(":
                pos += 1
                kids = [parse_subtree()]
                while tokens[pos] == ",":
                    pos += 1
                    kids.append(parse_subtree())
                assert tokens[pos] == ")", "unbalanced Newick"
                pos += 1
                if pos < len(tokens) and not tokens[pos].startswith(":") and tokens[pos] not in "(),;":
                    pos += 1                       # inner label / support value
                ln = None
                if pos < len(tokens) and tokens[pos].startswith(":"):
                    ln = float(tokens[pos][1:]); pos += 1
                return ("inner", kids, ln)
            name = tokens[pos]; pos += 1
            ln = None
            if pos < len(tokens) and tokens[pos].startswith(":"):
                ln = float(tokens[pos][1:]); pos += 1
            return ("tip", name_to_num[name], ln)

        def build(sub) -> Node:
            """Materialise a subtree; returns the record that must be hooked to the parent."""
            kind = sub[0]
            if kind == "tip":
                return t.nodep[sub[1]]
            kids = sub[1]
            assert len(kids) == 2, "only binary trees are supported"
            n = t.nodep[nxt[0]]; nxt[0] += 1
            for slot, kid in zip((n.next, n.next.next), kids):
                hookup(slot, build(kid), zlist(kid[2]))
            return n

        root = parse_subtree()
        assert root[0] == "inner"
        kids = root[1]
        if len(kids) == 3:
            n = t.nodep[nxt[0]]; nxt[0] += 1
            for slot, kid in zip((n, n.next, n.next.next), kids):
                hookup(slot, build(kid), zlist(kid[2]))
        elif len(kids) == 2:                       # rooted: join the two root children
            a, b = build(kids[0]), build(kids[1])
            la = (kids[0][2] or 0.0) + (kids[1][2] or 0.0) if (kids[0][2] is not None or kids[1][2] is not None) else None
            hookup(a, b, zlist(la))
        else:
            raise ValueError("root must have 2 or 3 children")
        assert nxt[0] == 2 * t.mxtips - 1, "tree does not contain every taxon exactly once"
        t.start = t.nodep[1]
        return t

    def to_newick(self) -> str:
        """Unrooted Newick without branch lengths, rooted at start.back (cf. Tree2String treeIO.c:324)."""
        def rec(p: Node) -> str:
            if self.isTip(p.number):
                return self.names[p.number - 1]
            return "(" + rec(p.next.back) + "," + rec(p.next.next.back) + ")"
        p = self.start.back
        return "(" + self.names[self.start.number - 1] + "," + rec(p.next.back) + "," + rec(p.next.next.back) + ");"

    # ---- traversal descriptors ---------------------------------------------------------------
    def computeTraversalInfo(self, p: Node, ti: List[Entry], partialTraversal: bool) -> None:
        """Append, in post-order, the CLVs below record p that must be (re)computed
        (newviewGenericSpecial.c:691-811).  Iterative to survive deep caterpillar trees."""
        if self.isTip(p.number):
            return
        # frames: (record, stage); stage 0 = descend, 1 = emit
        stack = [(p, 0)]
        while stack:
            node, stage = stack.pop()
            q, r = node.next.back, node.next.next.back
            qt, rt = self.isTip(q.number), self.isTip(r.number)
            if stage == 0:
                stack.append((node, 1))
                if qt and rt:
                    continue
                if qt or rt:
                    inner = q if rt else r
                    if (not inner.x) or (not partialTraversal):
                        stack.append((inner, 0))
                else:
                    # the reference recurses into q first, then r: push r first so q pops first
                    if (not r.x) or (not partialTraversal):
                        stack.append((r, 0))
                    if (not q.x) or (not partialTraversal):
                        stack.append((q, 0))
                continue
            if not node.x:
                self.getxnode(node)
            if qt and rt:
                case = TIP_TIP
            elif qt or rt:
                case = TIP_INNER
                if rt:                    # the tip always goes into q (newviewGenericSpecial.c:742-750)
                    q, r = r, q
            else:
                case = INNER_INNER
            ti.append(Entry(case, node.number, q.number, r.number, list(q.z), list(r.z)))

    def full_traversal(self, p: Optional[Node] = None) -> List[Entry]:
        """Descriptor of evaluateGeneric(tr, p, fullTraversal=TRUE) (evaluateGenericSpecial.c:897-953):
        entry 0 is the branch (p, p.back), followed by every inner CLV in post-order."""
        p = p or self.start
        assert self.isTip(p.number)
        q = p.back
        ti = [Entry(-1, p.number, q.number, 0, list(q.z), [0.0] * self.numBranches)]
        self.computeTraversalInfo(q, ti, False)
        return ti

    def partial_traversal(self, p: Node) -> List[Entry]:
        """Descriptor of evaluateGeneric(tr, p, FALSE) / makenewzGeneric: entry 0 = branch (p, p.back),
        then whatever is not oriented towards that branch."""
        q = p.back
        ti = [Entry(-1, p.number, q.number, 0, list(q.z), [0.0] * self.numBranches)]
        if not self.isTip(p.number) and not p.x:
            self.computeTraversalInfo(p, ti, True)
        if not self.isTip(q.number) and not q.x:
            self.computeTraversalInfo(q, ti, True)
        return ti

    def newview_traversal(self, p: Node) -> List[Entry]:
        """Descriptor of newviewGeneric(tr, p, masked) (newviewGenericSpecial.c:1523-1556)."""
        ti: List[Entry] = []
        self.computeTraversalInfo(p, ti, True)
        return ti

    def tip_case_counts(self, ti: Sequence[Entry]):
        c = [0, 0, 0]
        for e in ti:
            if e.tipCase >= 0:
                c[e.tipCase] += 1
        return c
