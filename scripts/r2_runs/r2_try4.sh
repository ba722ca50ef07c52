#!/bin/bash
# producer/consumer warp split of the persistent FP32 filter: 2+6 (192 sources per tile), 3+5 (160), 4+4 (128)
cp vif_b200/lib/libvif_b200_qxmatch.so /tmp/orig.so
echo "c192 pb4 (shipped): $(python scripts/r2_var.py)"
for v in c160_pb4:160 c160_pb6:160 c128_pb4:128 c192_pb3:192; do
  f=${v%%:*}; t=${v##*:}
  cp vif_b200/lib/variant_$f.so vif_b200/lib/libvif_b200_qxmatch.so
  echo "$f: $(XM_NN32_TILE=$t python scripts/r2_var.py)"
done
cp /tmp/orig.so vif_b200/lib/libvif_b200_qxmatch.so
