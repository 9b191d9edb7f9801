#!/bin/bash
# round 2, call 9 (1 GPU): score_insertions parity tests (call 8 stopped before them), the rewritten Newton-Raphson kernel
# (state in shared memory, 2 instead of 4 grid barriers per iteration, one CTA and no copies for a handful of tiles),
# small-call latency and search times
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "score_insertions or footprint or newton or root_invariance or lg4_mixed or psr_cat" 2>&1 | tail -6 | tee gpurun_out/c9_tests_a.txt
python tools/gpu_latency.py 2>&1 | tee gpurun_out/c9_latency.txt
python - <<'PY' 2>&1 | tee gpurun_out/c9_search_time.txt
import sys, os, time, tempfile, subprocess
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from dropin_util import write_inputs, run_mode, REFERENCE, small_dna_case
for case in ("dna49", "synthetic 60 taxa x 20000 sites"):
    with tempfile.TemporaryDirectory() as tmp:
        if case == "dna49": b, t, _ = write_inputs("dna49", tmp)
        else: b, t = small_dna_case(tmp, taxa=60, sites=20000, seed=3)
        for name, exe, env in (("drop-in", "bin/examl-b200", {}), ("drop-in + batched insertions", "bin/examl-b200-spr", {"EXAML_B200_SPR_STATS": "1"})):
            t0 = time.time()
            r, best, _ = run_mode(os.path.abspath(exe), b, t, tmp, name.replace(" ", "_").replace("+", "p"), ["-m", "GAMMA", "-f", "d"], env=env, timeout=1200)
            print(case, "|", name, "| -f d wall %.1f s" % (time.time() - t0), "best", best, [l for l in r.stdout.splitlines() if "B200 back-end" in l])
PY
timeout 600 python -m pytest tests/test_gpu_dropin.py -x -q 2>&1 | tail -3 | tee gpurun_out/c9_tests_b.txt
