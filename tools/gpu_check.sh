#!/bin/bash
# Run on the GPU box (via gpurun): parity tests, then a tile-size sweep of the DNA bench.
# usage: tools/gpu_check.sh [tiles...]
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for t in "$@"; do
  python bench.py --tile "$t" --steps 20 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_tile_$t.json
  python - "$t" <<'PY'
import json, sys
t = sys.argv[1]
d = json.load(open(f"gpurun_out/bench_tile_{t}.json"))
print("tile", t, "value %.0f M/s" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"],
      "frac %.3f" % d["roofline"]["frac"], d["clocks"])
PY
done
