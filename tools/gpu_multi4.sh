#!/bin/bash
# usage: tools/gpu_multi4.sh N  -- correctness of both all-reduce paths at N ranks, then the bench on both
N=${1:-4}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py 2>&1 | grep MULTI_GPU_CHECK | cut -c1-120
for mode in p2p nccl; do
  if [ $mode = p2p ]; then export EXB_P2P=1; else unset EXB_P2P; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 40 2>/dev/null | tail -1 > gpurun_out/scale_${mode}_$N.json
  python - $N $mode <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/scale_{sys.argv[2]}_{sys.argv[1]}.json"))
print("gpus", sys.argv[1], sys.argv[2], "value %.0f M/s" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "lnl", d["config"]["lnl"], "|", d["config"].get("allreduce"))
PY
done
