#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for cfg in "64 3" "64 2" "128 3" "128 2" "256 2"; do
  set -- $cfg
  EXB_DNA_MINB=$2 python bench.py --tile $1 --steps 20 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/b.json
  python - "$cfg" <<'PY'
import json, sys
d = json.load(open("gpurun_out/b.json"))
print("tile/minb", sys.argv[1], "value %.0f M/s" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "frac %.3f" % d["roofline"]["frac"])
PY
done
