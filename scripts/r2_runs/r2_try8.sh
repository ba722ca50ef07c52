#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | grep -v "^frame" | tail -2
echo "deferred   : $(timeout 120 python scripts/r2_var.py)"
echo "round-1 ord: $(XM_NN32_NODEFER=1 timeout 120 python scripts/r2_var.py)"
echo "deferred   : $(timeout 120 python scripts/r2_var.py)"
timeout 300 python -m pytest tests/test_gpu_fullsize.py -x -q -k "c2" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2t8_bench.json 2> gpurun_out/r2t8_bench.err; cut -c1-180 gpurun_out/r2t8_bench.json
