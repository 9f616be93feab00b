#!/bin/bash
# final sanity of round 1: smoke() and the mixed-precision suite (its solve path gained the transposed branch)
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r34_smoke.log
timeout 40 python -m pytest tests/test_gpu_mixed.py -q -k "not full_size" 2>&1 | tail -6 | tee gpurun_out/r34_mixed.log
