#!/bin/bash
echo "base         : $(python scripts/r2_var.py)"
echo "dbg=1 (no candidate loads)   : $(XM_NN32_DBG=1 python scripts/r2_var.py)"
echo "dbg=2 (consumers skip compute): $(XM_NN32_DBG=2 python scripts/r2_var.py)"
