#!/bin/bash
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $R --master-port 29511 scripts/dist_check.py > gpurun_out/r2_n2_dist_full.log 2>&1; grep "OK\|FAIL\|Error\|error" gpurun_out/r2_n2_dist_full.log | head -30
XM_DIST_TIMING=1 timeout 400 $R --master-port 29515 bench.py --gpus 2 --config c5 --rows 200000000 --rows2 20000000 --steps 2 --warmup 3 --no-e2e > gpurun_out/r2_n2_c5_t.log 2>&1; grep "banded\|metric" gpurun_out/r2_n2_c5_t.log | tail -3 | cut -c1-400
XM_C5_HOME=1 XM_DIST_TIMING=1 timeout 400 $R --master-port 29516 bench.py --gpus 2 --config c5 --rows 200000000 --rows2 20000000 --steps 2 --warmup 3 --no-e2e > gpurun_out/r2_n2_c5_home.log 2>&1; grep "banded\|metric" gpurun_out/r2_n2_c5_home.log | tail -3 | cut -c1-400
