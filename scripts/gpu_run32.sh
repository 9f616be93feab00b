#!/bin/bash
# last GPU call of round 1: the shared-memory batched class (65..160 rows) with the widen and batched suites
timeout 80 python -m pytest tests/test_gpu_widen.py tests/test_gpu_batched.py -q 2>&1 | tail -25 | tee gpurun_out/r32_tests.log
