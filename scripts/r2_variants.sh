#!/bin/bash
# occupancy experiment: the same library built with different __launch_bounds__ of ring_nn32_kernel
cp vif_b200/lib/libvif_b200_qxmatch.so /tmp/orig.so
for b in 4 5 6; do
  cp vif_b200/lib/variant_minb$b.so vif_b200/lib/libvif_b200_qxmatch.so
  for t in 192 224 256; do
    echo "minb $b tile $t: $(XM_NN32_TILE=$t python scripts/profile_target.py 10000000 5 2>&1 | tail -1 | grep -o "'ms_forward.*rows_refined_rev': [0-9]*")"
  done
done
cp /tmp/orig.so vif_b200/lib/libvif_b200_qxmatch.so
