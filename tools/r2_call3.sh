#!/bin/bash
# round 2, call 3: DNA kernel v7 (d-table rows in shared memory, cp.async prefetch ring): full GPU tests, C3 / C5 timings,
# ncu --set full of the three shipped traverse kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for ns in 4 2; do
  EXB_DNA_NS=$ns python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c3 ns $ns: value %.0f ms/step %.3f phases %s parity %s' % (d['value'], d['ms_per_step'], d['kernel_phases_ms'], d['parity']['rel_err']))"
done
for ns in 4 2 1; do
  EXB_DNA_NS=$ns python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c5 ns $ns: value %.0f ms/step %.3f phases %s parity %s' % (d['value'], d['ms_per_step'], d['kernel_phases_ms'], d['parity']['rel_err']))"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:traverse_dna -s 3 -c 1 -f -o gpurun_out/prof_dna_v7 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_dna_v7.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:traverse_aa -s 3 -c 1 -f -o gpurun_out/prof_aa_v7 \
    python bench.py --workload aa --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_aa_v7.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:traverse_dna -s 3 -c 1 -f -o gpurun_out/prof_c5_v1 \
    python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c5_v1.log 2>&1
ls -la gpurun_out | grep "prof_.*v[17]"
