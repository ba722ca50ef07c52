#!/bin/bash
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 200 $R --master-port 29511 scripts/dist_check.py binned,ties_across_shards,banded,partitioned > gpurun_out/r2_n4_dist.log 2>&1; grep "^OK\|FAIL\|Error" gpurun_out/r2_n4_dist.log | head -12
XM_BENCH_NO_C5=1 timeout 200 $R --master-port 29512 bench.py --gpus 4 --steps 10 --warmup 3 --no-e2e > gpurun_out/r2_n4_c2.log 2>&1; grep metric gpurun_out/r2_n4_c2.log | cut -c1-200; grep "Error" gpurun_out/r2_n4_c2.log | head -3
