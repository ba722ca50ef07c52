# round-end single-GPU pass: tests, bench (c2 default, c5), ncu launch list of the bench command
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/n1_tests.log
timeout 600 python bench.py > gpurun_out/n1_bench_c2.json 2> gpurun_out/n1_bench_c2.err
timeout 600 python bench.py --config c5 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/n1_bench_c5.json 2> gpurun_out/n1_bench_c5.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/n1_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/n1_ncu.log 2>&1
cat gpurun_out/n1_tests.log
