#!/bin/bash
# dynamic tile scheduling of the persistent filter: parity, then A/B of (tile assignment) x (what the reverse launch waits for)
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | grep -v "^frame" | tail -2
echo "dyn, after kernel : $(XM_NN32_AFTER=kernel timeout 120 python scripts/r2_var.py)"
echo "dyn, after lane   : $(timeout 120 python scripts/r2_var.py)"
echo "static, kernel    : $(XM_NN32_DYN=0 XM_NN32_AFTER=kernel timeout 120 python scripts/r2_var.py)"
echo "static, lane      : $(XM_NN32_DYN=0 timeout 120 python scripts/r2_var.py)"
echo "dyn, concurrent   : $(XM_NN32_CONCURRENT=1 timeout 120 python scripts/r2_var.py)"
timeout 300 python -m pytest tests/test_gpu_fullsize.py -x -q -k "c2" 2>&1 | tail -2
