#!/bin/bash
# round 2, call 2: new DNA kernel (tip LUT, carry flags), AA permuted fragments as the only body, C4/C5 tests, default bench with secondary
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
./tools/microbench/write_floor 2>&1 | tee gpurun_out/write_floor.txt
( time python bench.py --steps 20 --warmup 3 ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -3 gpurun_out/bench_default.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
def show(tag, r):
    print(tag, "value %.0f ms/step %.3f kernel %.3f phases %s roofline frac %.3f parity %s" % (
        r["value"], r["ms_per_step"], r["kernel_ms_per_step"], {k: round(v, 4) for k, v in r["kernel_phases_ms"].items()},
        r["roofline"]["frac"], r.get("parity", {}).get("rel_err")))
show("C3", d)
for k, r in d.get("secondary", {}).items():
    if "error" in r: print(k, r)
    else: show(k, r); print("   ", r.get("step_breakdown_ms"), r.get("cpu_baseline", {}).get("value"))
print("C3 cpu", d.get("cpu_baseline", {}).get("value"), "e2e", d["e2e"]["value"], "clocks", d["clocks"])
PY
EXB_DNA_NO_LUT=1 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('no-LUT: value %.0f ms/step %.3f kernel_ms %.3f' % (d['value'], d['ms_per_step'], d['kernel_ms_per_step']))"
