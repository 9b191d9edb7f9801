#!/bin/bash
# last GPU seconds of the round: the specialised DNA entry variant (parity, then speed), then the workflow tests
export EXB_DNA_SPEC=1
timeout 40 python -m pytest tests/test_gpu_parity.py -x -q -k "reference_goldens or clv_scaler or scaling_path_iid or large_tree" 2>&1 | tail -2
timeout 45 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('SPEC value %.0f ms/step %.3f kernel_ms %.3f lnl %r' % (d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['config']['lnl']))"
unset EXB_DNA_SPEC
timeout 40 python -m pytest tests/test_gpu_dropin.py -x -q -k "workflows" 2>&1 | tail -2
