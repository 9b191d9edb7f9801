#!/bin/bash
# round 2, call 12 (1 GPU): one-CTA Newton-Raphson with thread-per-site sweeps; AA kernel with 16 warps per CTA (128
# registers, spills) against the shipped 12; ncu --set full with source of the AA kernel on a reduced C4 (40 k sites)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "newton or root_invariance or lg4 or goldens" 2>&1 | tail -3 | tee gpurun_out/c12_tests_a.txt
python tools/gpu_latency.py 2>&1 | grep -E "makenewz|evaluate, 3|newview, 3" | tee gpurun_out/c12_latency.txt
python - <<'PY' 2>&1 | tee gpurun_out/c12_search_time.txt
import sys, os, time, tempfile, subprocess
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from dropin_util import write_inputs, run_mode
with tempfile.TemporaryDirectory() as tmp:
    b, t, _ = write_inputs("dna49", tmp)
    for name, exe in (("drop-in", "bin/examl-b200"), ("drop-in + batched insertions", "bin/examl-b200-spr")):
        t0 = time.time()
        r, best, _ = run_mode(os.path.abspath(exe), b, t, tmp, name.replace(" ", "_").replace("+", "p"), ["-m", "GAMMA", "-f", "d"], timeout=1200)
        print("dna49 |", name, "| -f d wall %.1f s" % (time.time() - t0), "best", best)
PY
for lib in libexaml_b200.so libexaml_b200_w16.so; do
EXB_LIBRARY=$PWD/examl_b200/$lib timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "goldens and aa" 2>&1 | tail -1
EXB_LIBRARY=$PWD/examl_b200/$lib python bench.py --workload aa --steps 8 --warmup 3 --no-cpu-baseline --no-secondary 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('aa $lib: value %.0f ms/step %.3f phases %s parity %s' % (d['value'], d['ms_per_step'], d['kernel_phases_ms'], d['parity']['rel_err']))" | tee -a gpurun_out/c12_aa_warps.txt
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:traverse_aa -s 3 -c 1 -f -o gpurun_out/prof_aa_v8 \
    python bench.py --workload aa --sites 40000 --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_aa_v8.log 2>&1
ls -la gpurun_out | grep prof_aa_v8
