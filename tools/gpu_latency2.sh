#!/bin/bash
python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -3
python tools/gpu_latency.py
timeout 600 python tools/gpu_e2e_dropin.py 50 60000 d 2>&1 | grep -v reference | head -3
