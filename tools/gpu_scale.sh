#!/bin/bash
# usage: tools/gpu_scale.sh N...   (under gpurun --gpus max(N)): the driver's own launch line per N
mkdir -p gpurun_out
for n in "$@"; do
  if [ $n -eq 1 ]; then python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/scale_err_$n.log | tail -1 > gpurun_out/scale_$n.json
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 2>gpurun_out/scale_err_$n.log | tail -1 > gpurun_out/scale_$n.json; fi
  python - $n <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/scale_{n}.json"))
    print("gpus", n, "value %.0f M/s" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "lnl", d["config"]["lnl"], d["clocks"])
except Exception as ex:
    print("gpus", n, "FAILED", ex); print(open(f"gpurun_out/scale_err_{n}.log").read()[-1500:])
PY
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
