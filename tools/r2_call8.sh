#!/bin/bash
# round 2, call 8 (1 GPU): batched insertion scoring (exb_score_insertions) -- parity tests, drop-in search through the
# patched rearrangeBIG, search timing against the unpatched drop-in and the reference; SEV tests that call 7 did not reach;
# C3 / C4 timing with the scoring branch in the traverse kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "score_insertions or sev" 2>&1 | tail -8 | tee gpurun_out/c8_tests_a.txt
timeout 900 python -m pytest tests/test_gpu_dropin.py -x -q -s 2>&1 | grep -v "^$" | tail -14 | tee gpurun_out/c8_tests_b.txt
python - <<'PY' 2>&1 | tee gpurun_out/c8_search_time.txt
import sys, os, time, tempfile, subprocess
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from dropin_util import write_inputs, run_mode, REFERENCE, small_dna_case
for case in ("dna49", "synthetic 60 taxa x 20000 sites"):
    with tempfile.TemporaryDirectory() as tmp:
        if case == "dna49": b, t, _ = write_inputs("dna49", tmp)
        else: b, t = small_dna_case(tmp, taxa=60, sites=20000, seed=3)
        for name, exe, env in (("drop-in", "bin/examl-b200", {}), ("drop-in + batched insertions", "bin/examl-b200-spr", {"EXAML_B200_SPR_STATS": "1"}),
                               ("reference AVX 1 core", REFERENCE, {})):
            t0 = time.time()
            r, best, _ = run_mode(os.path.abspath(exe), b, t, tmp, name.replace(" ", "_").replace("+", "p"), ["-m", "GAMMA", "-f", "d"], env=env, timeout=1200)
            print(case, "|", name, "| -f d wall %.1f s" % (time.time() - t0), "best", best, [l for l in r.stdout.splitlines() if "batched insertion" in l])
PY
for w in dna aa; do
python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-secondary 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w: value %.0f ms/step %.3f phases %s parity %s' % (d['value'], d['ms_per_step'], d['kernel_phases_ms'], d['parity']['rel_err']))" | tee -a gpurun_out/c8_bench.txt
done
