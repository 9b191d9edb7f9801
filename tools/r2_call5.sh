#!/bin/bash
# round 2, call 5 (2 GPUs): tcgen05 building-block self-test; per-rank timing of both all-reduce paths, alternating
mkdir -p gpurun_out
timeout 120 ./tools/microbench/umma_tf32 2>&1 | tee gpurun_out/umma_tf32.txt
for mode in nccl p2p nccl p2p; do
  if [ $mode = nccl ]; then export EXB_P2P=0; else unset EXB_P2P; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 40 --no-secondary 2>/dev/null | tail -1 > gpurun_out/scale_${mode}_2.json
  python - $mode <<'PY'
import json, sys
mode = sys.argv[1]
d = json.load(open(f"gpurun_out/scale_{mode}_2.json"))
print("gpus 2", mode, "value %.0f M/s" % d["value"], "ms/step %.3f" % d["ms_per_step"], "kernel %.3f" % d["kernel_ms_per_step"], {k: round(v, 3) for k, v in d["kernel_phases_ms"].items()},
      "e2e %.0f" % d["e2e"]["value"], "clocks", d["clocks"])
print("   per rank [wall, wall_e2e, kernels, tables, traverse, finish]:", d.get("per_rank_ms_per_step", {}).get("ranks"))
PY
done
