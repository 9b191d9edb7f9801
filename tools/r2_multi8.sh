#!/bin/bash
# round 2, 8-GPU call: correctness of the in-kernel all-reduce / device Newton-Raphson on 8 ranks, then the bench
# (default = in-kernel all-reduce with all secondaries; N = 4 and the NCCL path for comparison)
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py 2>&1 | grep MULTI_GPU_CHECK | cut -c1-400
show() {
python - $1 <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(sys.argv[1], "gpus", d["n_gpus"], "value %.0f M/s" % d["value"], "ms/step %.3f" % d["ms_per_step"], "kernel %.3f" % d["kernel_ms_per_step"], {k: round(v, 3) for k, v in d["kernel_phases_ms"].items()},
      "e2e %.0f" % d["e2e"]["value"], "|", d["config"].get("allreduce"), d["clocks"])
print("   per rank [wall, wall_e2e, kernels, tables, traverse, finish]:", d.get("per_rank_ms_per_step", {}).get("ranks"))
for k, r in d.get("secondary", {}).items():
    if "error" in r: print("  ", k, r)
    else: print("  ", k, "value %.0f ms/step %.3f kernel %.3f finish %.3f e2e %.0f parity %s | %s" % (r["value"], r["ms_per_step"], r["kernel_ms_per_step"], r["finish_ms"], r["e2e"]["value"], r.get("parity", {}).get("rel_err"), r.get("limiter")))
PY
}
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 60 2>gpurun_out/scale8.err | tail -1 > gpurun_out/scale8_p2p.json
tail -2 gpurun_out/scale8.err; show gpurun_out/scale8_p2p.json
EXB_P2P=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 60 --no-secondary 2>/dev/null | tail -1 > gpurun_out/scale8_nccl.json
show gpurun_out/scale8_nccl.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 60 --no-secondary 2>/dev/null | tail -1 > gpurun_out/scale4_p2p.json
show gpurun_out/scale4_p2p.json
python bench.py --steps 40 --no-secondary --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/scale1.json
show gpurun_out/scale1.json
