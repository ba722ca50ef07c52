#!/bin/bash
# parity first, then variant timings of the pipelined index kernels and the producers' L2 prefetch
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -k "c2" 2>&1 | tail -3
echo "base (pipe=3, ctas=8, pf=1): $(python scripts/r2_var.py)"
for p in 0 1 2; do echo "XM_INDEX_PIPE=$p: $(XM_INDEX_PIPE=$p python scripts/r2_var.py)"; done
for c in 3 4 6 12; do echo "XM_INDEX_PIPE_CTAS=$c: $(XM_INDEX_PIPE_CTAS=$c python scripts/r2_var.py)"; done
echo "XM_NN32_PF=0: $(XM_NN32_PF=0 python scripts/r2_var.py)"
echo "XM_NN32_PERSIST=0: $(XM_NN32_PERSIST=0 python scripts/r2_var.py)"
python scripts/profile_target.py 10000000 3 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2t1_launches.csv python scripts/profile_target.py 10000000 3 > gpurun_out/r2t1_ncu.log 2>&1
python scripts/launch_shares.py gpurun_out/r2t1_launches.csv
