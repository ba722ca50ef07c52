#!/bin/bash
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "promised or row_ranges or errors" 2>&1 | tail -3
timeout 500 $R --master-port 29511 scripts/dist_check.py > gpurun_out/r2_n2_dist.log 2>&1; grep "OK\|FAIL\|Error\|error" gpurun_out/r2_n2_dist.log | head -40
XM_BENCH_NO_C5=1 XM_DIST_TIMING=1 timeout 300 $R --master-port 29513 bench.py --gpus 2 --steps 4 --warmup 3 --no-e2e > gpurun_out/r2_n2_c2_t.log 2>&1; grep sharded gpurun_out/r2_n2_c2_t.log | tail -2
timeout 600 $R --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_n2_bench.log 2>&1; grep metric gpurun_out/r2_n2_bench.log | cut -c1-300; tail -3 gpurun_out/r2_n2_bench.log | cut -c1-300
