#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_getrf.py -x -q -m gpu -k "single_rhs or solve_backward or streamed" 2>&1 | tail -15
timeout 200 python scripts/e2e_breakdown.py 8192 4096 16384 2>&1 | grep -v "^\[stream\]" | tail -14
