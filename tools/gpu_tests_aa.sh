#!/bin/bash
# full GPU test suite + one AA GAMMA bench line
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --workload aa --steps 10 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-200
