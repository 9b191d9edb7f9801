"""Insertion scan at scale: all lazy SPR insertions of pruned subtrees within a rearrangement radius, scored (a) by ONE
exb_score_insertions call per pruned subtree and (b) the reference's way, candidate by candidate (newviewGeneric +
evaluateGeneric per insertion, tests/spr_util.reference_scan) through the same engine.  Synthetic DNA, GTR+GAMMA.
usage: tools/gpu_spr_bench.py [taxa] [sites] [radius] [prunings]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from examl_b200.capi import Engine
from examl_b200.likelihood import Likelihood
from examl_b200.phylo import hookup
from examl_b200.spr import build_scan, remove_node, score_scan
from spr_util import prunable_nodes, reference_scan
from synth import build_engine, random_case

taxa = int(sys.argv[1]) if len(sys.argv) > 1 else 200
sites = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
radius = int(sys.argv[3]) if len(sys.argv) > 3 else 5
nprune = int(sys.argv[4]) if len(sys.argv) > 4 else 6
tree, parts, models = random_case(taxa, [sites], [4], seed=17)
eng = build_engine(lambda *a, **k: Engine(*a, **k), tree, parts, models)
eng.set_timing(False)
lk = Likelihood(tree, eng, 1)
lk.evaluateGeneric(tree.start, True)
rng = np.random.default_rng(3)
nodes = prunable_nodes(tree)
tb = tr = 0.0
ncand = 0
worst = 0.0
for pick in rng.choice(len(nodes), size=nprune, replace=False):
    p = nodes[pick]
    saved = (p.next.back, list(p.next.z), p.next.next.back, list(p.next.next.z))
    p1, p2 = remove_node(tree, p, [a * b for a, b in zip(p.next.z, p.next.next.z)])
    lk.newviewGeneric(p1); lk.newviewGeneric(p2)
    ti, cands = build_scan(tree, p, p1, p2, 1, radius)
    score_scan(eng, ti, len(cands), eng.max_entries())                 # warm-up (scratch nodes are allocated on first use)
    t0 = time.perf_counter()
    got = score_scan(eng, ti, len(cands), eng.max_entries())
    tb += time.perf_counter() - t0
    t0 = time.perf_counter()
    want = reference_scan(lk, p, p1, p2, 1, radius)
    tr += time.perf_counter() - t0
    ncand += len(cands)
    for k, (q, lh) in enumerate(want):
        worst = max(worst, abs(got[k][0] - lh[0]) / abs(lh[0]))
    hookup(p.next, saved[0], saved[1]); hookup(p.next.next, saved[2], saved[3])
    lk.newviewGeneric(p)
print("insertion scan, DNA %d taxa x %d sites, radius %d, %d pruned subtrees, %d candidates: batched %.1f ms (%.3f ms per candidate), "
      "candidate by candidate %.1f ms (%.3f ms per candidate): %.2fx; max rel. difference of the scores %.1e"
      % (taxa, sites, radius, nprune, ncand, 1e3 * tb, 1e3 * tb / ncand, 1e3 * tr, 1e3 * tr / ncand, tr / tb, worst))
