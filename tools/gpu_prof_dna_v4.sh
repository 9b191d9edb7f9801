#!/bin/bash
mkdir -p gpurun_out
timeout 100 ncu --set full --clock-control none --import-source on -k regex:traverse_dna -s 3 -c 1 -f -o gpurun_out/prof_dna_v4 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_dna_v4.log 2>&1
ls -la gpurun_out/ | grep v4
