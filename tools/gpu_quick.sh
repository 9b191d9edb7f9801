#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
python bench.py --workload aa --model gamma --steps 10 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_aa_gamma.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_aa_gamma.json"))
print("aa gamma value %.0f M/s" % d["value"], "ms/step %.3f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"])
PY
