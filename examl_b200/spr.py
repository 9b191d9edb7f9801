"""Batched scoring of lazy SPR insertions: the host side (Python mirror of the scan in examl_b200/csrc/axml_backend.c).

The reference tests ONE insertion at a time.  rearrangeBIG (searchAlgo.c:800) prunes the subtree at p
(removeNodeBIG searchAlgo.c:440), then addTraverseBIG (searchAlgo.c:785) walks outwards from the pruning point and
calls testInsertBIG (searchAlgo.c:682) on every branch between mintrav and maxtrav steps away: hook p into the branch
with both halves at sqrt(z) (insertBIG, non-thorough branch, searchAlgo.c:519-535), newviewGeneric(p),
evaluateGeneric(p->next->next), unhook.  Every evaluateGeneric re-orients the CLV on the near side of the candidate
branch, so walking outwards costs one re-orientation + one new-node CLV + one evaluation per candidate, each as its own
small call.

`build_scan` turns the whole walk into ONE mixed descriptor for Engine.scoreInsertions (exb_score_insertions):

  * the CLV that looks from the pruning point TOWARDS a candidate branch is computed level by level into scratch
    nodes (one per depth of the walk: a level's vector is needed by its own candidate and by both deeper levels, and
    is dead once the walk leaves the level), so the tree's own CLVs and x flags are not re-oriented at all;
  * subtrees hanging off the walk, the far end of a candidate and the pruned subtree must look towards the walk; where
    their x flag says they do not, the ordinary computeTraversalInfo entries are put in front (exactly what the
    reference's evaluateGeneric would do on demand) -- these are the only changes to the tree's state;
  * a scoring entry per candidate.

The order of the candidates is the order in which addTraverseBIG visits them, so a caller can replay the reference's
decisions (bestOfNode, endLH, the lhCutoff pruning of the walk) on the table of results.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

from .phylo import INNER_INNER, TIP_INNER, TIP_TIP, ZMAX, ZMIN, Entry, Node, Tree

SCORE = 4                 # EXB_SCORE
SCRATCH_NODES = 64        # EXB_SCRATCH_NODES


def _entry(tree: Tree, tip_case_bit: int, p_number: int, a_number: int, az: Sequence[float], b_number: int,
           bz: Sequence[float]) -> Entry:
    """A descriptor entry with children a and b; a tip child always goes into q (newviewGenericSpecial.c:742-750)."""
    at, bt = tree.isTip(a_number), tree.isTip(b_number)
    if bt and not at:
        a_number, az, b_number, bz = b_number, bz, a_number, az
        at, bt = bt, at
    case = TIP_TIP if (at and bt) else (TIP_INNER if at else INNER_INNER)
    return Entry(case | tip_case_bit, p_number, a_number, b_number, list(az), list(bz))


def lazy_insertion_z(z: Sequence[float]) -> List[float]:
    """Both halves of a branch a subtree is lazily inserted into (insertBIG searchAlgo.c:522-531)."""
    return [min(max(math.sqrt(v), ZMIN), ZMAX) for v in z]


def build_scan(tree: Tree, p: Node, p1: Node, p2: Node, mintrav: int, maxtrav: int) -> Tuple[List[Entry], List[Node]]:
    """Descriptor of all lazy insertions of the subtree pruned at p (p.back = the subtree; p1--p2 already joined, as after
    removeNodeBIG) that addTraverseBIG would test from p1 and from p2.  Returns (ti, candidates): ti[0] describes the
    subtree, candidates[k] is the record q of the k-th candidate branch (q, q.back) in visiting order.  Raises
    ValueError when the walk is deeper than the scratch nodes allow."""
    nb = tree.numBranches
    s = p.back
    ti: List[Entry] = [Entry(-1, s.number, s.number, 0, list(p.z[:nb]), [0.0] * nb)]
    cands: List[Node] = []
    scratch0 = 2 * tree.mxtips - 1

    def orient(rec: Node) -> None:
        """Make the CLV of rec's node look along rec (towards rec.back), like evaluateGeneric does on demand."""
        if not tree.isTip(rec.number) and not rec.x:
            tree.computeTraversalInfo(rec, ti, True)

    orient(s)

    def walk(q: Node, mintrav: int, maxtrav: int, depth: int, up_rec: Node, up_number: int) -> None:
        # candidate branch (q, r): r belongs to the ring of inner node R; up_rec is the record of that ring that looks
        # back at the pruning point, the third record carries the subtree hanging off the walk
        if depth >= SCRATCH_NODES:
            raise ValueError("rearrangement radius exceeds the scratch nodes of exb_score_insertions")
        r = q.back
        ring = [r, r.next, r.next.next]
        sib_rec = next(x for x in ring if x is not r and x is not up_rec)
        sib = sib_rec.back
        orient(sib)
        u_number = scratch0 + depth
        ti.append(_entry(tree, 0, u_number, up_number, up_rec.z[:nb], sib.number, sib_rec.z[:nb]))
        mintrav -= 1
        if mintrav <= 0:
            orient(q)
            zz = lazy_insertion_z(q.z[:nb])
            ti.append(_entry(tree, SCORE, len(cands), q.number, zz, u_number, zz))
            cands.append(q)
        maxtrav -= 1
        if not tree.isTip(q.number) and maxtrav > 0:
            walk(q.next.back, mintrav, maxtrav, depth + 1, q, u_number)
            walk(q.next.next.back, mintrav, maxtrav, depth + 1, q, u_number)

    for a, b in ((p1, p2), (p2, p1)):
        if tree.isTip(a.number):
            continue
        # a.back is b: the record `a` looks at the pruning point; b's CLV (or tip) must look at a
        orient(b)
        walk(a.next.back, mintrav, maxtrav, 0, a, b.number)
        walk(a.next.next.back, mintrav, maxtrav, 0, a, b.number)
    return ti, cands


def score_scan(engine, ti: List[Entry], ncand: int, max_entries: int, execute_model=None):
    """Run a scan through Engine.scoreInsertions in as many calls as the engine's descriptor capacity requires (scratch
    CLVs persist between calls, so the list may be cut anywhere).  Returns lnl[ncand, partitions]."""
    import numpy as np
    out = np.zeros((ncand, engine.nparts))
    body = ti[1:]
    at = 0
    while at < len(body) or (at == 0 and not body):
        chunk = body[at:at + max_entries - 1]
        # renumber this call's candidates from 0
        ids = [e.pNumber for e in chunk if e.tipCase >= 0 and (e.tipCase & SCORE)]
        remap = {k: i for i, k in enumerate(ids)}
        call = [ti[0]] + [Entry(e.tipCase, remap[e.pNumber], e.qNumber, e.rNumber, e.qz, e.rz)
                          if (e.tipCase & SCORE) else e for e in chunk]
        res = engine.scoreInsertions(call, len(ids), execute_model)
        for k, i in remap.items():
            out[k] = res[i]
        at += len(chunk)
        if not body:
            break
    return out


def remove_node(tree: Tree, p: Node, zqr: Sequence[float]) -> Tuple[Node, Node]:
    """Topology part of removeNodeBIG (searchAlgo.c:440-462): join p's two neighbours with branch value zqr (the
    reference optimises it with makenewzGeneric first), leave p.next / p.next.next dangling."""
    from .phylo import hookup
    q, r = p.next.back, p.next.next.back
    hookup(q, r, zqr)
    p.next.back = None
    p.next.next.back = None
    return q, r
