#!/bin/bash
# First GPU call of the next round (1 GPU, ~4 min): everything that could not be measured at the end of round 1.
#   1. full GPU test suite (includes test_many_partitions_shared_memory_constants, never run on a GPU yet)
#   2. DNA kernel sweep: body (EXB_DNA_SPEC 0 run-time selection | 1 specialised | 2 specialised + staged) x registers budget (EXB_DNA_MINB) x sites per thread (--tile 128 -> NS=2)
#   3. AA kernel: EXB_AA_PERM=1 parity + speed
#   4. ncu --set full of the specialised DNA body and of partial_kernel -> gpurun_out/prof_dna_v5 / prof_partial
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
# the staged DNA entry body (constant-outer / site-inner loops, EXB_DNA_SPEC=2) has never run on a GPU: parity first
EXB_DNA_SPEC=2 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
EXB_DNA_SPEC=2 EXB_DNA_MINB=3 python -m pytest tests/test_gpu_parity.py -x -q -k "goldens or clv_scaler or large_tree or many_partitions" 2>&1 | tail -2
for cfg in "1 2 256" "2 2 256" "2 3 256" "0 2 256" "1 3 256" "2 3 128" "2 2 128"; do
  set -- $cfg
  EXB_DNA_SPEC=$1 EXB_DNA_MINB=$2 python bench.py --steps 20 --warmup 3 --tile $3 --no-cpu-baseline --no-secondary 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('spec $1 minb $2 tile $3: value %.0f ms/step %.3f kernel_ms %.3f sm_mhz %s %s' % (d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons']))"
done
# AA: fragments with the permuted contraction index (EXB_AA_PERM=1: one shuffle per contraction instead of ten), never run on a GPU
EXB_AA_PERM=1 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
for perm in 0 1; do
  EXB_AA_PERM=$perm python bench.py --workload aa --steps 10 --warmup 3 --no-cpu-baseline --no-secondary 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('aa perm $perm: value %.0f ms/step %.3f kernel_ms %.3f' % (d['value'], d['ms_per_step'], d['kernel_ms_per_step']))"
done
timeout 120 ncu --set full --clock-control none --import-source on -k regex:traverse_dna -s 3 -c 1 -f -o gpurun_out/prof_dna_v5 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_dna_v5.log 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:partial_kernel -s 1 -c 1 -f -o gpurun_out/prof_partial \
    python tools/gpu_partial_bench.py dna > gpurun_out/ncu_partial.log 2>&1
ls -la gpurun_out | grep prof_
timeout 300 ncu --set full --clock-control none --import-source on -k regex:traverse_aa -s 3 -c 1 -f -o gpurun_out/prof_aa_v6 \
    python bench.py --workload aa --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_aa_v6.log 2>&1
ls -la gpurun_out | grep prof_
