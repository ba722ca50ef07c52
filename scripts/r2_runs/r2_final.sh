#!/bin/bash
# round-2 records on one B200: full GPU suite, the bench line, launch list, ncu summaries of the two filter kernels
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^frame" | tail -4 > gpurun_out/r2f_tests.txt; cat gpurun_out/r2f_tests.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; cut -c1-200 gpurun_out/r2f_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; cut -c1-200 gpurun_out/r2f_bench_ref.json
python scripts/profile_target.py 10000000 3 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches.csv python scripts/profile_target.py 10000000 3 > gpurun_out/r2f_ncu.log 2>&1
python scripts/launch_shares.py gpurun_out/r2f_launches.csv > gpurun_out/r2f_shares.txt 2>&1; cat gpurun_out/r2f_shares.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ring_nn32p -s 2 -c 1 -f -o gpurun_out/r2f_nn32p python scripts/profile_target.py 10000000 2 > gpurun_out/r2f_nn32p.log 2>&1
python scripts/profile_sphere.py > /dev/null 2>&1 && timeout 400 ncu --set full --clock-control none --import-source on -k regex:sphere_nn_filter -s 2 -c 1 -f -o gpurun_out/r2f_sphere python scripts/profile_sphere.py > gpurun_out/r2f_sphere.log 2>&1
ls -la gpurun_out/*.ncu-rep
