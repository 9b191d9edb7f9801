#!/bin/bash
# evaluatePartialGeneric on the GPU: parity tests, the PSR drop-in runs, and a whole-program -m PSR -f e timing
python -m pytest tests -m gpu -x -q -k "partial or psr" 2>&1 | tail -6
python tools/gpu_e2e_dropin.py 200 50000 e PSR 2>&1 | tail -4
