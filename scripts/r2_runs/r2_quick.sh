#!/bin/bash
# quick GPU check of the C2 hot path: parity tests, stage times, per-kernel launch list
python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
python scripts/profile_target.py 10000000 4 2>&1 | tail -1
for t in 256 224 192 160 128; do echo "tile $t"; XM_NN32_TILE=$t python scripts/profile_target.py 10000000 4 2>&1 | tail -1 | cut -c1-330; done
python scripts/profile_target.py 10000000 3 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_q_launches.csv python scripts/profile_target.py 10000000 3 > gpurun_out/r2_q_ncu.log 2>&1
python scripts/launch_shares.py gpurun_out/r2_q_launches.csv
