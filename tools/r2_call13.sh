#!/bin/bash
# round 2, call 13 (1 GPU): AA kernel -- one wave of equal tiles, idle warps of a partial chunk skip the tensor work, L2
# prefetch of the next entry's loaded children: parity + C4 timing; full GPU suite
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/c13_tests.txt
python bench.py --workload aa --steps 10 --warmup 3 --no-cpu-baseline --no-secondary 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('aa: value %.0f ms/step %.3f phases %s parity %s roofline frac %.3f' % (d['value'], d['ms_per_step'], d['kernel_phases_ms'], d['parity']['rel_err'], d['roofline']['frac']))" | tee gpurun_out/c13_aa.txt
python bench.py --workload aa --sites 25000 --steps 20 --warmup 3 --no-cpu-baseline --no-secondary 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('aa 25k sites (one GPU share of C4 at 8): value %.0f ms/step %.3f phases %s' % (d['value'], d['ms_per_step'], d['kernel_phases_ms']))" | tee -a gpurun_out/c13_aa.txt
