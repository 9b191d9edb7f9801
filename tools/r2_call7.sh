#!/bin/bash
# round 2, call 7 (1 GPU): state of HEAD after the SEV commit -- full GPU tests, default bench line, small-call latency
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/c7_tests.txt
( time timeout 600 python bench.py --steps 20 --warmup 3 ) > gpurun_out/c7_bench_default.json 2> gpurun_out/c7_bench_default.err
tail -4 gpurun_out/c7_bench_default.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c7_bench_default.json").read().strip().splitlines()[-1])
def show(tag, r):
    print(tag, "value %.0f ms/step %.3f kernel %.3f phases %s roofline %s parity %s" % (
        r["value"], r["ms_per_step"], r["kernel_ms_per_step"], {k: round(v, 4) for k, v in r["kernel_phases_ms"].items()},
        r["roofline"], r.get("parity", {}).get("rel_err")))
show("C3", d)
for k, r in d.get("secondary", {}).items():
    if "error" in r: print(k, r)
    else: show(k, r); print("   ", r.get("step_breakdown_ms"), r.get("cpu_baseline", {}).get("value"), r.get("limiter"))
print("C3 cpu", d.get("cpu_baseline"), "e2e", d["e2e"], "clocks", d["clocks"])
PY
timeout 300 bash tools/gpu_search_time.sh 2>&1 | tee gpurun_out/c7_latency.txt
