"""The reference's own way of scoring lazy SPR insertions, restated on the Python host mirror: the checker of the batched
scan (examl_b200/spr.py, exb_score_insertions)."""
from __future__ import annotations

from examl_b200.phylo import hookup
from examl_b200.spr import lazy_insertion_z


def reference_scan(lk, p, p1, p2, mintrav, maxtrav):
    """addTraverseBIG (searchAlgo.c:785) + testInsertBIG (searchAlgo.c:682), non-thorough, without the lhCutoff pruning:
    for every candidate hook p in (insertBIG searchAlgo.c:519-535), newviewGeneric(p), evaluateGeneric(p->next->next),
    unhook.  Returns [(q record, per-partition lnL)] in visiting order.  `lk` is a examl_b200.likelihood.Likelihood."""
    tree = lk.tr
    out = []

    def test_insert(q):
        r = q.back
        qz = list(q.z)
        zz = lazy_insertion_z(q.z[:tree.numBranches])
        hookup(p.next, q, zz)
        hookup(p.next.next, r, zz)
        lk.newviewGeneric(p)
        lk.evaluateGeneric(p.next.next, False)
        out.append((q, lk.perPartitionLH.copy()))
        hookup(q, r, qz)
        p.next.next.back = None
        p.next.back = None

    def add_traverse(q, mintrav, maxtrav):
        mintrav -= 1
        if mintrav <= 0:
            test_insert(q)
        maxtrav -= 1
        if not tree.isTip(q.number) and maxtrav > 0:
            add_traverse(q.next.back, mintrav, maxtrav)
            add_traverse(q.next.next.back, mintrav, maxtrav)

    for a in (p1, p2):
        if not tree.isTip(a.number):
            add_traverse(a.next.back, mintrav, maxtrav)
            add_traverse(a.next.next.back, mintrav, maxtrav)
    return out


def prunable_nodes(tree):
    """Inner-node records p whose subtree p.back can be pruned the way rearrangeBIG does on its p side (searchAlgo.c:812-819)."""
    out = []
    for i in range(tree.mxtips + 1, 2 * tree.mxtips - 1):
        n = tree.nodep[i]
        for p in (n, n.next, n.next.next):
            p1, p2 = p.next.back, p.next.next.back
            if not tree.isTip(p1.number) or not tree.isTip(p2.number):
                out.append(p)
    return out
