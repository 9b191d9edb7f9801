#!/bin/bash
python tools/gpu_e2e_dropin.py 200 400000 e PSR 2>&1 | tail -4
