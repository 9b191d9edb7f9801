#!/bin/bash
timeout 50 python -m pytest tests/test_gpu_dropin.py -x -q -k "workflows" 2>&1 | tail -4
