#!/bin/bash
# source-level ncu captures of the three heaviest kernels of the C2 step (one GPU)
python scripts/profile_target.py 10000000 2 > /dev/null 2>&1 || exit 1
for k in ring_nn32p place_kernel count_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r2_src_$k python scripts/profile_target.py 10000000 2 > gpurun_out/r2_src_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
