#!/bin/bash
# one 8-GPU box: correctness subset, the driver's bench line (C2 weak + c5_strong), per-phase timing of both
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "promised" > gpurun_out/r2_promised.log 2>&1; grep -v "^frame" gpurun_out/r2_promised.log | tail -2
timeout 300 $R --master-port 29511 scripts/dist_check.py binned,ties_across_shards,banded,banded_dups,partitioned > gpurun_out/r2_n8_dist.log 2>&1; grep "^OK\|FAIL\|Error" gpurun_out/r2_n8_dist.log | head -12
timeout 500 $R --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_n8_bench.log 2>&1; grep metric gpurun_out/r2_n8_bench.log | cut -c1-240; grep "Error" gpurun_out/r2_n8_bench.log | head -3
XM_BENCH_NO_C5=1 XM_DIST_TIMING=1 timeout 200 $R --master-port 29513 bench.py --gpus 8 --steps 4 --warmup 3 --no-e2e > gpurun_out/r2_n8_c2_t.log 2>&1; grep sharded gpurun_out/r2_n8_c2_t.log | tail -2
XM_DIST_TIMING=1 timeout 300 $R --master-port 29515 bench.py --gpus 8 --config c5 --steps 3 --warmup 3 --no-e2e > gpurun_out/r2_n8_c5_t.log 2>&1; grep "banded\|metric" gpurun_out/r2_n8_c5_t.log | tail -3 | cut -c1-330
