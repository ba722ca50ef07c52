#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | grep -v "^frame" | tail -2
echo "cursor placement : $(timeout 120 python scripts/r2_var.py)"
echo "rank array (r1)  : $(XM_INDEX_RANK=1 timeout 120 python scripts/r2_var.py)"
echo "cursor placement : $(timeout 120 python scripts/r2_var.py)"
python scripts/profile_target.py 10000000 3 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2t7_launches.csv python scripts/profile_target.py 10000000 3 > gpurun_out/r2t7_ncu.log 2>&1
python scripts/launch_shares.py gpurun_out/r2t7_launches.csv | head -8
