"""exb_evaluate_partial throughput: every site of the bench workload at one rate each (= one round of the PSR rate
scan of optimizeRateCategories), kernel time from the engine's CUDA events and wall time through the C ABI; the
oracle's restatement of evaluatePartialGeneric on one host core beside it.   usage: tools/gpu_partial_bench.py [dna|aa]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench
from examl_b200.capi import Engine
from examl_b200.phylo import Tree

wl = sys.argv[1] if len(sys.argv) > 1 else "dna"
taxa, sites, states = (200, 1_000_000, 4) if wl == "dna" else (500, 200_000, 20)
tree = Tree.random(taxa, seed=bench.TREE_SEED)
ti = tree.full_traversal()
tips = bench.make_sites(tree, 0, sites, wl)
wgt = np.ones(sites, dtype=np.int32)
eng = Engine(taxa, 1, 1)
eng.set_rate_model(0, 25)
eng.set_partition(0, states, wgt, tips)
eng.commit()
eng.set_model(0, bench.model_for(wl))
rates, cat = bench.psr_categories(sites)
eng.set_categories(0, rates, cat)
mo = np.zeros(sites, dtype=np.int32); si = np.arange(sites, dtype=np.int32); ra = rates[cat] * 0.9
tp = eng.pack(ti)
for _ in range(2):
    v = eng.evaluatePartial(tp, mo, si, ra)
ks, ws = [], []
for _ in range(5):
    t0 = time.perf_counter(); v = eng.evaluatePartial(tp, mo, si, ra); ws.append(time.perf_counter() - t0)
    ks.append(eng.last_kernel_ms() * 1e-3)
k, w = min(ks), min(ws)
upd = sites * (taxa - 2)
print(f"{wl}: {sites} (site, rate) pairs x {taxa - 2} entries: kernel {k * 1e3:.2f} ms = {sites / k / 1e6:.1f} M pairs/s "
      f"= {upd / k / 1e9:.2f} G site-updates/s; through the C ABI (host arrays in, values out) {w * 1e3:.2f} ms")
# the oracle (one core) on a sample
from oracle.oracle import OracleEngine
n = 2000
cpu = OracleEngine(taxa, 1, 1)
cpu.set_rate_model(0, 25); cpu.set_partition(0, states, wgt[:n], tips[:, :n]); cpu.commit(); cpu.set_model(0, bench.model_for(wl))
cpu.set_categories(0, rates, cat[:n])
t0 = time.perf_counter(); c = cpu.evaluatePartial(ti, mo[:n], si[:n], ra[:n]); dt = time.perf_counter() - t0
print(f"oracle port, 1 core: {n} pairs in {dt:.2f} s = {n / dt / 1e3:.1f} K pairs/s; max |gpu - oracle| rel "
      f"{np.max(np.abs(v[:n] - c) / np.maximum(1, np.abs(c))):.2e}")
