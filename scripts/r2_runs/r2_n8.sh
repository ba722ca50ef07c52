#!/bin/bash
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $R --master-port 29511 scripts/dist_check.py 2>&1 | grep "OK\|FAIL\|Error\|error" > gpurun_out/r2_n8_dist.log; tail -16 gpurun_out/r2_n8_dist.log
timeout 400 $R --master-port 29514 bench.py --gpus 8 --config c5 --steps 5 --warmup 3 --no-e2e > gpurun_out/r2_n8_c5.log 2>&1; grep metric gpurun_out/r2_n8_c5.log | cut -c1-260
XM_DIST_TIMING=1 timeout 400 $R --master-port 29515 bench.py --gpus 8 --config c5 --steps 2 --warmup 3 --no-e2e > gpurun_out/r2_n8_c5_t.log 2>&1; grep banded gpurun_out/r2_n8_c5_t.log | tail -2
XM_C5_HOME=1 timeout 400 $R --master-port 29516 bench.py --gpus 8 --config c5 --steps 5 --warmup 3 --no-e2e > gpurun_out/r2_n8_c5_home.log 2>&1; grep metric gpurun_out/r2_n8_c5_home.log | cut -c1-260
timeout 300 python bench.py --gpus 1 --config c5 --steps 5 --warmup 3 --no-e2e > gpurun_out/r2_n1_c5.log 2>&1; grep metric gpurun_out/r2_n1_c5.log | cut -c1-260
timeout 300 $R --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 --no-e2e > gpurun_out/r2_n8_c2_weak.log 2>&1; grep metric gpurun_out/r2_n8_c2_weak.log | cut -c1-260
XM_DIST_TIMING=1 timeout 300 $R --master-port 29513 bench.py --gpus 8 --steps 4 --warmup 3 --no-e2e > gpurun_out/r2_n8_c2_weak_t.log 2>&1; grep sharded gpurun_out/r2_n8_c2_weak_t.log | tail -2
