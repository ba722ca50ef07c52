#!/bin/bash
cp vif_b200/lib/libvif_b200_qxmatch.so /tmp/orig.so
run() { python scripts/profile_target.py 10000000 5 2>&1 | tail -1 | grep -o "'ms_forward.*rows_refined_rev': [0-9]*"; }
cp vif_b200/lib/variant_c192.so vif_b200/lib/libvif_b200_qxmatch.so
echo "c192 serial:     $(run)"
echo "c192 concurrent: $(XM_NN32_CONCURRENT=1 run)"
cp vif_b200/lib/variant_c160.so vif_b200/lib/libvif_b200_qxmatch.so
echo "c160 serial:     $(XM_NN32_TILE=160 run)"
cp /tmp/orig.so vif_b200/lib/libvif_b200_qxmatch.so
