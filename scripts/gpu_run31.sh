#!/bin/bash
# late round-1 check: the "widen" tests + the suites that cover code touched late in the round, then the nb sweep
timeout 100 python -m pytest tests/test_gpu_widen.py tests/test_gpu_batched.py tests/test_gpu_interface.py -q 2>&1 | tail -8 | tee gpurun_out/r31_tests.log
timeout 60 python scripts/sweep_nb.py 4096 8192 2>&1 | tee gpurun_out/r31_sweep_nb.log
