#!/bin/bash
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^frame" | tail -4 > gpurun_out/r2h_tests.txt; cat gpurun_out/r2h_tests.txt
timeout 400 $R --master-port 29511 scripts/dist_check.py > gpurun_out/r2h_n2_dist.log 2>&1; grep -c "^OK" gpurun_out/r2h_n2_dist.log; grep "FAIL\|Error\|error" gpurun_out/r2h_n2_dist.log | head -5
XM_BENCH_NO_C5=1 timeout 300 $R --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2h_n2_c2.log 2>&1; grep metric gpurun_out/r2h_n2_c2.log | cut -c1-200
python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; cut -c1-200 gpurun_out/r2h_bench.json
python scripts/profile_target.py 10000000 3 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_launches.csv python scripts/profile_target.py 10000000 3 > gpurun_out/r2h_ncu.log 2>&1
python scripts/launch_shares.py gpurun_out/r2h_launches.csv > gpurun_out/r2h_shares.txt 2>&1; cat gpurun_out/r2h_shares.txt
