#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
for a in 0 296 592 888 1480; do echo "ahead $a: $(XM_NN32_AHEAD=$a python scripts/profile_target.py 10000000 5 2>&1 | tail -1 | grep -o "'ms_forward.*rows_refined_rev': [0-9]*")"; done
