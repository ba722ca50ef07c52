#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | grep -v "^frame" | tail -3
timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -k "c2" 2>&1 | tail -2
echo "column index : $(python scripts/r2_var.py)"
echo "count/place  : $(XM_INDEX_COL=0 python scripts/r2_var.py)"
python scripts/profile_target.py 10000000 3 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2t5_launches.csv python scripts/profile_target.py 10000000 3 > gpurun_out/r2t5_ncu.log 2>&1
python scripts/launch_shares.py gpurun_out/r2t5_launches.csv
