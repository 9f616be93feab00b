#!/bin/bash
B200LU_STREAM_DBG=1 timeout 200 python scripts/e2e_breakdown.py 8192 2>&1 | tail -40
for c in 4 16; do echo "== chunks $c"; B200LU_H2D_CHUNKS=$c timeout 200 python scripts/e2e_breakdown.py 8192 2>&1 | grep "stream_h2d=1" | tail -1; done
timeout 200 python scripts/e2e_breakdown.py 16384 4096 2>&1 | tail -10
