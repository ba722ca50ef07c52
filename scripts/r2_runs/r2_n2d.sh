#!/bin/bash
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "promised or row_ranges or errors" > gpurun_out/r2_promised.log 2>&1; grep -v "^frame" gpurun_out/r2_promised.log | tail -4
timeout 500 $R --master-port 29511 scripts/dist_check.py > gpurun_out/r2_n2_dist.log 2>&1; grep -c "^OK" gpurun_out/r2_n2_dist.log; grep "FAIL\|Error\|error" gpurun_out/r2_n2_dist.log | head -10
XM_BENCH_NO_C5=1 XM_DIST_TIMING=1 timeout 300 $R --master-port 29513 bench.py --gpus 2 --steps 4 --warmup 3 --no-e2e > gpurun_out/r2_n2_c2_t.log 2>&1; grep sharded gpurun_out/r2_n2_c2_t.log | tail -2
XM_BENCH_NO_C5=1 timeout 300 $R --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e > gpurun_out/r2_n2_c2.log 2>&1; grep metric gpurun_out/r2_n2_c2.log | cut -c1-200; grep "Error" gpurun_out/r2_n2_c2.log | head -3
XM_BENCH_NO_C5=1 XM_C2_REPLICATE=1 timeout 300 $R --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e > gpurun_out/r2_n2_c2_repl.log 2>&1; grep metric gpurun_out/r2_n2_c2_repl.log | cut -c1-200
XM_BENCH_NO_C5=1 XM_DIST_EARLY_BCAST=1 XM_C2_REPLICATE=1 timeout 300 $R --master-port 29515 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e > gpurun_out/r2_n2_c2_r1flow.log 2>&1; grep metric gpurun_out/r2_n2_c2_r1flow.log | cut -c1-200
