#!/bin/bash
# the device residual check (B200LU_OPT_KEEP_A + b200lu_residual_norms) and late timings
timeout 50 python -m pytest tests/test_gpu_widen.py tests/test_gpu_interface.py -q -k "residual" 2>&1 | tail -25 | tee gpurun_out/r33_tests.log
timeout 40 python scripts/late_timings.py 2>&1 | tee gpurun_out/r33_timings.log
