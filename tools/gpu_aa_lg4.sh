#!/bin/bash
python -m pytest tests -m gpu -x -q -k "lg4 or aa or golden" 2>&1 | tail -3
python bench.py --workload aa --steps 10 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-200
python tools/gpu_partial_bench.py dna
python tools/gpu_partial_bench.py aa
