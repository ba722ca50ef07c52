#!/bin/bash
# state of the round-2 tree on one B200: GPU tests, persistent vs one-CTA-per-tile filter, bench line, launch list
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2s_tests.txt
for p in 1 0; do echo "persist $p: $(XM_NN32_PERSIST=$p python scripts/profile_target.py 10000000 6 2>&1 | tail -1)"; done > gpurun_out/r2s_persist.txt 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
python scripts/profile_target.py 10000000 3 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2s_launches.csv python scripts/profile_target.py 10000000 3 > gpurun_out/r2s_ncu.log 2>&1
python scripts/launch_shares.py gpurun_out/r2s_launches.csv > gpurun_out/r2s_shares.txt 2>&1
cat gpurun_out/r2s_tests.txt gpurun_out/r2s_persist.txt gpurun_out/r2s_shares.txt
