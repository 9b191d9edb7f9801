#!/bin/bash
# round 2, the 8-GPU call: in-kernel all-reduce / device Newton-Raphson on 8 ranks (tests/multi_gpu_check.py), then the default
# bench line at N=8 exactly as the driver launches it (C3 strong-scaled + secondary: C4 and the FULL C5, 10 000 partitions)
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py 2>&1 | grep MULTI_GPU_CHECK | cut -c1-400 | tee gpurun_out/multi8_check.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 2>gpurun_out/scale8.err | tail -1 > gpurun_out/scale8_p2p.json
tail -3 gpurun_out/scale8.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/scale8_p2p.json"))
print("gpus", d["n_gpus"], "value %.0f M/s" % d["value"], "ms/step %.3f" % d["ms_per_step"], "kernel %.3f" % d["kernel_ms_per_step"], {k: round(v, 3) for k, v in d["kernel_phases_ms"].items()},
      "e2e %.0f" % d["e2e"]["value"], "|", d["config"].get("allreduce"), d["clocks"])
print("   per rank [wall, wall_e2e, kernels, tables, traverse, finish]:", d.get("per_rank_ms_per_step", {}).get("ranks"))
for k, r in d.get("secondary", {}).items():
    if "error" in r: print("  ", k, r)
    else: print("  ", k, "value %.0f ms/step %.3f kernel %.3f e2e %.0f parity %s | %s" % (r["value"], r["ms_per_step"], r["kernel_ms_per_step"], r["e2e"]["value"], r.get("parity", {}).get("rel_err"), r.get("limiter")))
PY
