#!/bin/bash
for cfg in "4 8" "6 8" "4 16" "6 16"; do set -- $cfg; echo "== NEAR=$1 CS=$2"; B200LU_TRSV_DBG=1 B200LU_TRSV3_NEAR=$1 B200LU_TRSV3_CS=$2 timeout 120 python scripts/trsv_bench.py 8192 2>&1 | grep -v "mode=2" | tail -12; done
echo "== sizes, default"; timeout 120 python scripts/trsv_bench.py 2048 4096 16384 2>&1 | tail -6
