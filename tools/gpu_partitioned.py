"""C5-shaped workload at reduced scale: many DNA partitions x 1 000 sites on a 354-taxon tree (BASELINE config 5 is
10 k partitions on 8 GPUs = 1 250 partitions per GPU).  Checks parity against the oracle on a few partitions and
reports throughput of a full-traversal evaluateGeneric.   usage: tools/gpu_partitioned.py [partitions] [sites_per_partition]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from examl_b200 import data as exdata
from examl_b200.capi import Engine
from examl_b200.phylo import Tree
from oracle.oracle import OracleEngine
from synth import random_models

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1250
W = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
T = 354
tree = Tree.random(T, seed=5)
ti = tree.full_traversal()
t0 = time.time()
models = random_models([4] * 8, seed=3)
tips = exdata.synthetic_dna(tree, W * 8, seed=11, mode="signal")
eng = Engine(T, P, 1)
for m in range(P):
    blk = tips[:, (m % 8) * W:(m % 8 + 1) * W]
    eng.set_partition(m, 4, np.ones(W, dtype=np.int32), np.ascontiguousarray(blk))
eng.commit()
for m in range(P):
    eng.set_model(m, models[m % 8])
print("setup %.1f s" % (time.time() - t0))
out = eng.evaluateIterative(ti).copy()
# parity on the 8 distinct (alignment block, model) combinations
orc = OracleEngine(T, 8, 1)
for m in range(8):
    orc.set_partition(m, 4, np.ones(W, dtype=np.int32), np.ascontiguousarray(tips[:, m * W:(m + 1) * W]))
    orc.set_model(m, models[m])
want = orc.evaluateIterative(ti)
err = max(abs(out[m] - want[m % 8]) / abs(want[m % 8]) for m in range(P))
print("max rel err vs oracle over %d partitions: %.2e" % (P, err))
n = 5
t0 = time.perf_counter(); dev = 0.0
for _ in range(n):
    eng.evaluateIterative(ti, out=out); dev += eng.last_kernel_ms()
wall = (time.perf_counter() - t0) / n
upd = (T - 2) * P * W
print("partitions %d x %d sites, %d taxa: %.2f ms/step wall, %.2f ms kernels -> %.0f M site-updates/s (kernels), %.0f (wall)"
      % (P, W, T, wall * 1e3, dev / n, upd / (dev / n) / 1e3, upd / wall / 1e6))
mask = np.ones(P, dtype=np.uint8); mask[::2] = 0
t0 = time.perf_counter(); eng.evaluateIterative(ti, mask, out=out); print("half masked: %.2f ms" % ((time.perf_counter() - t0) * 1e3))
