#!/bin/bash
mkdir -p gpurun_out
for cfg in "dna psr" "aa psr" "aa gamma"; do
  set -- $cfg
  python bench.py --workload $1 --model $2 --steps 10 2>/dev/null | tail -1 > gpurun_out/bench_$1_$2.json
  python - $1 $2 <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/bench_{sys.argv[1]}_{sys.argv[2]}.json"))
print(sys.argv[1], sys.argv[2], "value %.0f M/s" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"],
      "frac %.3f" % d["roofline"]["frac"], "cpu", d.get("cpu_baseline", {}).get("value"), d.get("cpu_baseline", {}).get("cores"))
PY
done
