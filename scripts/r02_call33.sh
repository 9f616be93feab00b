#!/bin/bash
# Round 2, GPU call 33 (one B200): racecheck of the cluster panel kernel after the parity double-buffering
mkdir -p gpurun_out
timeout 170 compute-sanitizer --tool racecheck --print-limit 40 --log-file gpurun_out/r02_sanitizer_racecheck_panel_after.txt \
    python scripts/sanitize_driver.py panel1 > gpurun_out/r02_sanitizer_racecheck_panel_after.out 2>&1
echo "== racecheck panel1: exit $?"; tail -3 gpurun_out/r02_sanitizer_racecheck_panel_after.out | cut -c1-200
grep -E "RACECHECK SUMMARY|Race reported" gpurun_out/r02_sanitizer_racecheck_panel_after.txt | sed 's/0x[0-9a-f]*//g' | sort | uniq -c | sort -rn | head -12 | cut -c1-260
