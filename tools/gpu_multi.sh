#!/bin/bash
# usage: tools/gpu_multi.sh N   (run under gpurun --gpus N): correctness of both all-reduce paths and of the device-resident
# Newton-Raphson across ranks, then the bench at N GPUs (default = in-kernel all-reduce, EXB_P2P=0 = NCCL)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -$N
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py 2>&1 | grep -v "^W0\|^\*\*\*" | tail -4 | cut -c1-600
EXB_P2P=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tests/multi_gpu_check.py 2>&1 | grep -v "^W0\|^\*\*\*" | tail -3 | cut -c1-300
for mode in ${MODES:-p2p nccl}; do
  if [ $mode = nccl ]; then export EXB_P2P=0; else unset EXB_P2P; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 40 $2 2> gpurun_out/scale_${mode}_$N.err | tail -1 > gpurun_out/scale_${mode}_$N.json
  tail -2 gpurun_out/scale_${mode}_$N.err
  python - $N $mode <<'PY'
import json, sys
n, mode = sys.argv[1], sys.argv[2]
d = json.load(open(f"gpurun_out/scale_{mode}_{n}.json"))
print("gpus", n, mode, "value %.0f M/s" % d["value"], "ms/step %.3f" % d["ms_per_step"], "kernel %.3f" % d["kernel_ms_per_step"], d["kernel_phases_ms"],
      "e2e %.0f" % d["e2e"]["value"], "lnl", d["config"]["lnl"], "|", d["config"].get("allreduce"))
print("   per rank [wall, wall_e2e, kernels, tables, traverse, finish]:", d.get("per_rank_ms_per_step", {}).get("ranks"))
for k, r in d.get("secondary", {}).items():
    if "error" in r: print("  ", k, r)
    else: print("  ", k, "value %.0f ms/step %.3f kernel %.3f finish %.3f e2e %.0f parity %s | %s" % (r["value"], r["ms_per_step"], r["kernel_ms_per_step"], r["finish_ms"], r["e2e"]["value"], r.get("parity", {}).get("rel_err"), r["config"]["allreduce"]))
PY
done
unset EXB_P2P
timeout 600 python -m pytest tests/test_gpu_dropin.py -x -q -k two_ranks 2>&1 | tail -2
