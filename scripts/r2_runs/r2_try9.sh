#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | grep -v "^frame" | tail -2
XM_INDEX_RANK=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "sphere or full_sky or bucket or promised or golden or seeded" 2>&1 | grep -v "^frame" | tail -2
timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -k "c5" 2>&1 | tail -2
timeout 400 python bench.py --config c5 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2t9_c5.json 2> gpurun_out/r2t9_c5.err; cut -c1-170 gpurun_out/r2t9_c5.json
echo "c2: $(timeout 120 python scripts/r2_var.py)"
