#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --workload aa --steps 5 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_aa.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_aa.json"))
print("AA value %.0f M/s" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "frac %.3f" % d["roofline"]["frac"], d["clocks"])
PY
