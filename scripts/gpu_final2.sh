#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/final_bench_default.json 2>gpurun_out/final_bench_default.err; tail -1 gpurun_out/final_bench_default.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('default', round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']), round(d['e2e']['ms_per_step'],1), d.get('e2e_matrix_rhs'), round(d['cpu_baseline']['value']), d['roofline']['frac'], d['clocks']['samples'])"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2>/dev/null; cut -c1-200 gpurun_out/final_bench_reference.json
timeout 600 python bench.py --n 32768 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/final_bench_32768.json 2>/dev/null; tail -1 gpurun_out/final_bench_32768.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('n32768', round(d['value']), round(d['getrf_ms'],1), round(d['getrs_ms'],1), round(d['getrf_gflops']), d['roofline']['frac'], d['roofline']['getrf_frac_of_fp64_peak'])"
timeout 600 python bench.py --workload mixed --n 16384 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/final_bench_mixed_16384.json 2>/dev/null; tail -1 gpurun_out/final_bench_mixed_16384.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('mixed16384', round(d['value']), round(d['getrf_ms'],1), round(d['getrs_ms'],1), d['roofline']['achieved'], d['roofline']['frac'], d['backward_error'])"
timeout 600 python bench.py --workload batched > gpurun_out/final_bench_batched.json 2>/dev/null; tail -1 gpurun_out/final_bench_batched.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('batched', round(d['value']), d['ms_per_step'], d['roofline']['frac'])"
