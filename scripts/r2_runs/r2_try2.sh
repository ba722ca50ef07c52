#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -k "c2" 2>&1 | tail -2
echo "base (pf=0): $(python scripts/r2_var.py)"
echo "XM_NN32_PF=1: $(XM_NN32_PF=1 python scripts/r2_var.py)"
echo "base again : $(python scripts/r2_var.py)"
