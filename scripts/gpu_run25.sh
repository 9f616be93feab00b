#!/bin/bash
for cfg in "4 8" "4 16" "6 16"; do set -- $cfg; echo "== NEAR=$1 CS=$2"; B200LU_TRSV_DBG=1 B200LU_TRSV3_NEAR=$1 B200LU_TRSV3_CS=$2 timeout 120 python scripts/trsv_bench.py 8192 2>&1 | grep -v "mode=2" | grep "trsv_mode\|clusters\|CTA 0\|CTA 5" ; done
echo "== no dbg"; for cs in 8 16; do B200LU_TRSV3_CS=$cs timeout 120 python scripts/trsv_bench.py 4096 8192 16384 2>&1 | grep "mode=0"; done
