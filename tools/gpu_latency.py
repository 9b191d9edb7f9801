"""Latency of the small calls that dominate tree search (testData/49-sized: 628 patterns, 49 taxa)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from examl_b200.capi import Engine
from golden_util import load_engine, load_json, traversal_of

first = load_json("dna49.first.json")
eng = load_engine(lambda *a, **k: Engine(*a, **k), first)
ti = traversal_of(first)
eng.evaluateIterative(ti)
def bench(name, f, n=2000):
    for _ in range(50): f()
    t0 = time.perf_counter()
    for _ in range(n): f()
    dt = (time.perf_counter() - t0) / n
    print(f"{name:40s} {dt*1e6:8.1f} us/call   kernel-side {eng.last_kernel_ms()*1e3:7.1f} us")
bench("evaluate full traversal (47 newviews)", lambda: eng.evaluateIterative(ti))
bench("evaluate, 3 newviews", lambda: eng.evaluateIterative(ti[:1] + ti[-3:]))
bench("evaluate, 0 newviews", lambda: eng.evaluateIterative(ti[:1]))
bench("newview, 3 entries", lambda: eng.newviewIterative(ti[-3:]))
bench("makenewz prepare (sum), 1 newview", lambda: eng.makenewzIterative(ti[:1] + ti[-1:]))
bench("core (one NR iteration)", lambda: eng.execCore([np.log(0.8)]))
br = ti[:1] + ti[-1:]
z, ok, it = eng.makenewz(br, [0.9], 10)
bench("makenewz on the device (%d NR iterations), 1 newview" % it, lambda: eng.makenewz(br, [0.9], 10))
eng.set_timing(False)
bench("  ... without CUDA events: makenewz", lambda: eng.makenewz(br, [0.9], 10))
bench("  ... without CUDA events: evaluate, 3 newviews", lambda: eng.evaluateIterative(ti[:1] + ti[-3:]))
bench("  ... without CUDA events: newview, 3 entries", lambda: eng.newviewIterative(ti[-3:]))
