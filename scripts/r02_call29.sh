#!/bin/bash
# Round 2, GPU call 29 (one B200): compute-sanitizer over small instances of every device path, then the full suite
mkdir -p gpurun_out
run() {   # tool, seconds, driver argument
    timeout $2 compute-sanitizer --tool $1 --print-limit 30 --log-file gpurun_out/r02_sanitizer_$1.txt \
        python scripts/sanitize_driver.py $3 > gpurun_out/r02_sanitizer_$1.out 2>&1
    echo "== $1: exit $?"; tail -2 gpurun_out/r02_sanitizer_$1.out | cut -c1-200
    grep -E "ERROR SUMMARY|Invalid|Uninitialized|hazard|  at " gpurun_out/r02_sanitizer_$1.txt | sed 's/0x[0-9a-f]*//g' | sort | uniq -c | sort -rn | head -10 | cut -c1-240
}
run memcheck 200 all
run initcheck 200 all
run racecheck 150 batched
timeout 400 python -m pytest tests -q -m gpu -x > gpurun_out/r02_final_tests.log 2>&1; tail -3 gpurun_out/r02_final_tests.log | cut -c1-400
