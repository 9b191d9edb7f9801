#!/bin/bash
# usage: tools/gpu_multi.sh N   (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py 2>&1 | grep -v "^W0\|^\*\*\*" | tail -5
EXB_P2P=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tests/multi_gpu_check.py 2>&1 | grep -v "^W0\|^\*\*\*" | tail -3
for n in 1 $N; do
  if [ $n -eq 1 ]; then python bench.py --gpus 1 --steps 20 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/scale_$n.json
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 20 2>/dev/null | tail -1 > gpurun_out/scale_$n.json; fi
  python - $n <<'PY'
import json, sys
n = sys.argv[1]
d = json.load(open(f"gpurun_out/scale_{n}.json"))
print("gpus", n, "value %.0f M/s" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "lnl", d["config"]["lnl"], d["config"].get("allreduce"))
PY
done
EXB_P2P=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 20 2>/dev/null | tail -1 > gpurun_out/scale_p2p_$N.json
python - $N <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/scale_p2p_{sys.argv[1]}.json"))
print("gpus", sys.argv[1], "in-kernel all-reduce (EXB_P2P=1): value %.0f M/s" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], d["config"].get("allreduce"))
PY
timeout 600 python -m pytest tests/test_gpu_dropin.py -x -q -k two_ranks 2>&1 | tail -2
