"""Batched lazy-SPR scoring, host logic (CPU): the mixed descriptor of examl_b200/spr.py executed by the oracle's
orc_score_insertions equals the reference's candidate-by-candidate procedure (tests/spr_util.reference_scan), and the
tree's state after a scan is a consistent one."""
import numpy as np
import pytest

from examl_b200.likelihood import Likelihood
from examl_b200.spr import build_scan, remove_node, score_scan
from spr_util import prunable_nodes, reference_scan
from synth import build_engine, random_case


def _two(mxtips, widths, kinds, seed, nb=1, mode="signal"):
    """Two identical (tree, engine) pairs: one for the batched scan, one for the reference procedure."""
    out = []
    for _ in range(2):
        tree, parts, models = random_case(mxtips, widths, kinds, seed=seed, mode=mode, num_branches=nb)
        out.append((tree, parts, models))
    return out


@pytest.mark.parametrize("mxtips,widths,kinds,nb,mintrav,maxtrav", [
    (12, [60], [4], 1, 1, 3),
    (20, [40, 25], [4, 20], 1, 1, 5),
    (20, [30, 30, 20], [4, 4, 20], 3, 2, 6),          # -M: one branch slot per partition; mintrav > 1
    (9, [50], [20], 1, 1, 10),                        # radius larger than the tree
])
def test_batched_scan_equals_candidate_by_candidate(oracle_factory, mxtips, widths, kinds, nb, mintrav, maxtrav):
    (tree, parts, models), (tree2, _, _) = _two(mxtips, widths, kinds, seed=11, nb=nb)
    eng = build_engine(oracle_factory, tree, parts, models)
    eng2 = build_engine(oracle_factory, tree2, parts, models)
    lk, lk2 = Likelihood(tree, eng, len(parts)), Likelihood(tree2, eng2, len(parts))
    lk.evaluateGeneric(tree.start, True); lk2.evaluateGeneric(tree2.start, True)
    rng = np.random.default_rng(5)
    nodes = prunable_nodes(tree)
    picks = rng.choice(len(nodes), size=min(6, len(nodes)), replace=False)
    total = 0
    for pick in picks:
        p = nodes[pick]
        p_2 = [x for x in prunable_nodes(tree2)][pick]
        assert p.number == p_2.number
        saved = [(p.next.back, list(p.next.z)), (p.next.next.back, list(p.next.next.z))]
        saved2 = [(p_2.next.back, list(p_2.next.z)), (p_2.next.next.back, list(p_2.next.next.z))]
        zqr = [a * b for a, b in zip(p.next.z, p.next.next.z)]
        p1, p2 = remove_node(tree, p, zqr)
        q1, q2 = remove_node(tree2, p_2, zqr)
        # removeNodeBIG's makenewzGeneric(q, r) leaves both neighbours looking at each other (searchAlgo.c:452)
        lk2.newviewGeneric(q1); lk2.newviewGeneric(q2)
        ti, cands = build_scan(tree, p, p1, p2, mintrav, maxtrav)
        got = score_scan(eng, ti, len(cands), max_entries=mxtips + 1)
        want = reference_scan(lk2, p_2, q1, q2, mintrav, maxtrav)
        assert [c.number for c in cands] == [q.number for q, _ in want]
        total += len(cands)
        for k, (q, lh) in enumerate(want):
            assert np.all(np.abs(got[k] - lh) <= 1e-11 * np.abs(lh)), (k, got[k], lh)
        # put the subtree back where it was (rearrangeBIG searchAlgo.c:838-840) and check the tree still evaluates
        from examl_b200.phylo import hookup
        for t, pp, sv in ((tree, p, saved), (tree2, p_2, saved2)):
            hookup(pp.next, sv[0][0], sv[0][1]); hookup(pp.next.next, sv[1][0], sv[1][1])
        lk.newviewGeneric(p); lk2.newviewGeneric(p_2)
        a = lk.evaluateGeneric(p, False); b = lk2.evaluateGeneric(p_2, False)
        assert abs(a - b) <= 1e-11 * abs(b)
    assert total > 0
    a = lk.evaluateGeneric(tree.start, True)
    b = lk.evaluateGeneric(tree.nodep[mxtips + 2], False)
    assert abs(a - b) <= 1e-11 * abs(a)


def test_scan_cut_into_several_calls(oracle_factory):
    """Scratch CLVs persist between calls: a descriptor capacity of a few entries gives the same table."""
    (tree, parts, models), _ = _two(16, [80], [4], seed=3)
    eng = build_engine(oracle_factory, tree, parts, models)
    lk = Likelihood(tree, eng, 1)
    lk.evaluateGeneric(tree.start, True)
    p = prunable_nodes(tree)[7]
    p1, p2 = remove_node(tree, p, [a * b for a, b in zip(p.next.z, p.next.next.z)])
    ti, cands = build_scan(tree, p, p1, p2, 1, 6)
    whole = score_scan(eng, ti, len(cands), max_entries=10_000)
    pieces = score_scan(eng, ti, len(cands), max_entries=4)
    assert len(cands) >= 8 and np.array_equal(whole, pieces)
