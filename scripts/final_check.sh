timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/final_tests.log
XM_SPHERE_QUEUE=1 timeout 300 python -m pytest tests -m gpu -x -q -k "exact_grid or full_sky or random or sphere" 2>&1 | tail -3 > gpurun_out/final_tests_queue.log
XM_SPHERE_QUEUE=1 ITERS=2 timeout 300 python scripts/c5_probe.py 1e9 1e8 2>&1 | grep -v "^generated" > gpurun_out/final_c5_queue.log
cat gpurun_out/final_tests.log gpurun_out/final_tests_queue.log gpurun_out/final_c5_queue.log
