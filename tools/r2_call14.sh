#!/bin/bash
# round 2, call 14 (1 GPU): which of the AA changes of call 13 costs time -- one-wave tiling vs 288-site tiles, L2 prefetch on / off;
# insertion scan at scale (tools/gpu_spr_bench.py)
mkdir -p gpurun_out
for cfg in "1 1" "0 0" "1 0" "0 1"; do
  set -- $cfg
  EXB_AA_ONE_WAVE=$1 EXB_AA_PREFETCH=$2 python bench.py --workload aa --steps 8 --warmup 3 --no-cpu-baseline --no-secondary 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('aa one_wave=$1 prefetch=$2: ms/step %.3f traverse %.3f clocks %s' % (d['ms_per_step'], d['kernel_phases_ms']['traverse'], d['clocks']['sm_mhz']))" | tee -a gpurun_out/c14_aa.txt
done
python tools/gpu_spr_bench.py 200 100000 5 6 2>&1 | tail -1 | tee gpurun_out/c14_spr_bench.txt
python tools/gpu_spr_bench.py 100 2000 10 6 2>&1 | tail -1 | tee -a gpurun_out/c14_spr_bench.txt
