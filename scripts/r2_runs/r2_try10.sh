#!/bin/bash
cp vif_b200/lib/libvif_b200_qxmatch.so /tmp/orig.so
echo "4 CTAs/SM, 64 regs, pb3 (shipped): $(timeout 120 python scripts/r2_var.py)"
for f in pctas3 pctas3_pb6; do cp vif_b200/lib/variant_$f.so vif_b200/lib/libvif_b200_qxmatch.so; echo "$f: $(timeout 120 python scripts/r2_var.py)"; done
cp /tmp/orig.so vif_b200/lib/libvif_b200_qxmatch.so
