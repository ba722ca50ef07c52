#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
for p in 1 0; do echo "persist $p: $(XM_NN32_PERSIST=$p python scripts/profile_target.py 10000000 5 2>&1 | tail -1 | grep -o "'ms_total.*rows_refined_rev': [0-9]*")"; done
