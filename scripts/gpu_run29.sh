#!/bin/bash
timeout 120 python scripts/trsv_bench.py 8192 16384 2>&1 | grep "mode=0"
timeout 120 python -m pytest tests/test_gpu_getrf.py -x -q -m gpu -k "single_rhs or solve_backward" 2>&1 | tail -2
for c in 6 12 16; do echo "== chunks $c"; B200LU_H2D_CHUNKS=$c timeout 200 python scripts/e2e_breakdown.py 8192 2>&1 | grep "stream_h2d=1" | tail -1; done
