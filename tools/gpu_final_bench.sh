#!/bin/bash
mkdir -p gpurun_out
python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_dna.json
python bench.py --workload aa --steps 10 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_aa.json
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_ref.json
python - <<'PY'
import json
for f in ("bench_dna", "bench_aa", "bench_ref"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value %.0f" % d["value"], "ms/step %.3f" % d["ms_per_step"], "kernel ms", d.get("kernel_ms_per_step"), "e2e %.0f" % d["e2e"]["value"],
              "frac", d.get("roofline", {}).get("frac"), "dram_frac", d.get("roofline", {}).get("dram_frac"), "cpu", d.get("cpu_baseline", {}).get("value"),
              d.get("cpu_baseline", {}).get("cores"), "parity", d.get("parity", {}).get("rel_err"), "launches", d.get("gpu_launches"), d.get("clocks"))
    except Exception as ex:
        print(f, "FAILED", ex)
PY
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 24 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
grep -c traverse gpurun_out/launches_final.csv
