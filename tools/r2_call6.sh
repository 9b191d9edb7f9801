#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -s -k "tf32x3" 2>&1 | tail -15
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "goldens or c4_shape or lg4" 2>&1 | tail -3
EXB_AA_PRECISION=tf32x3 python bench.py --workload aa --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/aa_tc.err | tail -1 > gpurun_out/bench_aa_tc.json
tail -3 gpurun_out/aa_tc.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_aa_tc.json").read().strip().splitlines()[-1])
print("aa tf32x3: value %.0f ms/step %.3f phases %s parity %s" % (d["value"], d["ms_per_step"], d["kernel_phases_ms"], d.get("parity")))
PY
