#!/bin/bash
# round 2, call 3: C4/C5 shape tests, ncu --set full of the three shipped traverse kernels, C5 sites-per-thread sweep
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -k "c4_shape or c5_shape or many_partitions" 2>&1 | tail -3
for ns in 4 2 1; do
  EXB_DNA_NS_MANY=$ns python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c5 ns $ns: value %.0f ms/step %.3f phases %s parity %s' % (d['value'], d['ms_per_step'], d['kernel_phases_ms'], d['parity']['rel_err']))"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:traverse_dna -s 3 -c 1 -f -o gpurun_out/prof_dna_v6 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_dna_v6.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:traverse_aa -s 3 -c 1 -f -o gpurun_out/prof_aa_v7 \
    python bench.py --workload aa --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_aa_v7.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:traverse_dna -s 3 -c 1 -f -o gpurun_out/prof_c5_v1 \
    python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c5_v1.log 2>&1
ls -la gpurun_out | grep "prof_.*v[167]"
