#!/bin/bash
# 8-GPU call of the next round (~2 min run, charged x8): both all-reduce paths, correctness first
mkdir -p gpurun_out
for mode in nccl p2p; do
  if [ $mode = p2p ]; then export EXB_P2P=1; else unset EXB_P2P; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py 2>&1 | grep MULTI_GPU_CHECK | cut -c1-100
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 60 2>/dev/null | tail -1 > gpurun_out/scale8_$mode.json
  python - $mode <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/scale8_{sys.argv[1]}.json"))
print("8 gpus", sys.argv[1], "value %.0f M/s" % d["value"], "ms/step %.3f" % d["ms_per_step"], "kernel ms %.3f" % d["kernel_ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "|", d["config"].get("allreduce"))
PY
done
