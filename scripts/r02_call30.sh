#!/bin/bash
# Round 2, GPU call 30 (one B200): memcheck over every device path with the mailbox panel kernel (see sanitize_driver.py)
mkdir -p gpurun_out
SAN_PANEL_MODE=1 timeout 400 compute-sanitizer --tool memcheck --print-limit 30 --log-file gpurun_out/r02_sanitizer_memcheck_mailbox.txt \
    python scripts/sanitize_driver.py all > gpurun_out/r02_sanitizer_memcheck_mailbox.out 2>&1
echo "== memcheck (panel mode 1): exit $?"; tail -4 gpurun_out/r02_sanitizer_memcheck_mailbox.out | cut -c1-200
grep -E "ERROR SUMMARY|Invalid|  at " gpurun_out/r02_sanitizer_memcheck_mailbox.txt | sed 's/0x[0-9a-f]*//g' | sort | uniq -c | sort -rn | head -10 | cut -c1-240
