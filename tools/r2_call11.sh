#!/bin/bash
# round 2, call 11 (1 GPU): the one-CTA Newton-Raphson kernel with everything on chip -- parity, latency, search times
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "newton or root_invariance or lg4 or goldens or masks" 2>&1 | tail -4 | tee gpurun_out/c11_tests_a.txt
python tools/gpu_latency.py 2>&1 | tee gpurun_out/c11_latency.txt
python - <<'PY' 2>&1 | tee gpurun_out/c11_search_time.txt
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
PY
timeout 900 python -m pytest tests/test_gpu_dropin.py -x -q 2>&1 | tail -3 | tee gpurun_out/c11_tests_b.txt
