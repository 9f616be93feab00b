#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(sorted(d.keys())); print(round(d['value']), round(d['ms_per_step'],2), d.get('e2e'), d.get('cpu_baseline'), d['gpu_launches'], d['clocks'], d['roofline']['frac'])"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-300
timeout 300 python bench.py --workload batched 2>&1 | tail -1 | cut -c1-250
