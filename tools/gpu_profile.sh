#!/bin/bash
# usage: tools/gpu_profile.sh <tag> <kernel-regex> [bench args...]
# one `ncu --set full` capture of the named kernel (launch index 3 = after warm-up) + launch list
tag=$1; shift; kre=$1; shift
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/ncu_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:$kre -s 3 -c 1 -f -o gpurun_out/prof_$tag \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline "$@" >> gpurun_out/ncu_$tag.log 2>&1
ls -la gpurun_out/
