#!/bin/bash
# round 2, call 10 (1 GPU): small-call path (descriptor in the kernel parameters, finish folded into the traverse kernel),
# one-CTA Newton-Raphson with the sum table in shared memory: full GPU test suite, latency, search times
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/c10_tests.txt
python tools/gpu_latency.py 2>&1 | tee gpurun_out/c10_latency.txt
EXB_NO_SMALL_CALLS=1 python tools/gpu_latency.py 2>&1 | grep "without" | sed 's/^/[EXB_NO_SMALL_CALLS=1] /' | tee -a gpurun_out/c10_latency.txt
python - <<'PY' 2>&1 | tee gpurun_out/c10_search_time.txt
import sys, os, time, tempfile, subprocess
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from dropin_util import write_inputs, run_mode, REFERENCE, small_dna_case
for case in ("dna49", "synthetic 60 taxa x 20000 sites"):
    with tempfile.TemporaryDirectory() as tmp:
        if case == "dna49": b, t, _ = write_inputs("dna49", tmp)
        else: b, t = small_dna_case(tmp, taxa=60, sites=20000, seed=3)
        for name, exe, env in (("drop-in", "bin/examl-b200", {}), ("drop-in + batched insertions", "bin/examl-b200-spr", {})):
            t0 = time.time()
            r, best, _ = run_mode(os.path.abspath(exe), b, t, tmp, name.replace(" ", "_").replace("+", "p"), ["-m", "GAMMA", "-f", "d"], env=env, timeout=1200)
            print(case, "|", name, "| -f d wall %.1f s" % (time.time() - t0), "best", best)
    if case == "dna49":
        with tempfile.TemporaryDirectory() as tmp:
            b, t, _ = write_inputs("dna49", tmp)
            for name, exe in (("drop-in", "bin/examl-b200"), ("reference AVX 1 core", REFERENCE)):
                t0 = time.time()
                r, best, _ = run_mode(os.path.abspath(exe), b, t, tmp, "E" + name[:3].replace("-", ""), ["-m", "GAMMA", "-f", "e"], timeout=1200)
                print(case, "|", name, "| -f e wall %.1f s" % (time.time() - t0))
PY
