#!/bin/bash
python tools/gpu_two_ranks.py
python tools/gpu_two_ranks.py EXB_P2P=1
timeout 900 python -m pytest tests/test_gpu_dropin.py -x -q -k "two_ranks or host_newton" 2>&1 | tail -3
