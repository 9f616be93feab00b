#!/bin/bash
for cfg in "4 8" "4 16"; do set -- $cfg; echo "== NEAR=$1 CS=$2, chain does not wait for far sums"; B200LU_TRSV_DBGFLAGS=1 B200LU_TRSV_DBG=1 B200LU_TRSV3_NEAR=$1 B200LU_TRSV3_CS=$2 timeout 120 python scripts/trsv_bench.py 8192 2>&1 | grep -v "mode=2" | grep "trsv_mode\|CTA 0\|CTA 5" ; done
