#!/bin/bash
# Round 2, GPU call 31 (one B200): racecheck (shared-memory hazards) over every fused panel class
mkdir -p gpurun_out
timeout 420 compute-sanitizer --tool racecheck --print-limit 40 --log-file gpurun_out/r02_sanitizer_racecheck_panels.txt \
    python scripts/sanitize_driver.py panels > gpurun_out/r02_sanitizer_racecheck_panels.out 2>&1
echo "== racecheck panels: exit $?"; tail -7 gpurun_out/r02_sanitizer_racecheck_panels.out | cut -c1-200
grep -E "RACECHECK SUMMARY|hazard|  at |Device Frame" gpurun_out/r02_sanitizer_racecheck_panels.txt | sed 's/0x[0-9a-f]*//g' | sort | uniq -c | sort -rn | head -16 | cut -c1-260
