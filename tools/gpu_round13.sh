#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_v7.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu_v7.log; tail -5 gpurun_out/r02_pytest_gpu_v7.log | cut -c1-300
run() { timeout 200 python bench.py --workload c1 --steps 2 --warmup 3 --lanes $1 --no-cpu-baseline 2>gpurun_out/r02_c1_l$1.err | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lanes $1:', round(d['lanes']['solves_per_s'],1), 'one at a time', round(d['one_at_a_time']['e2e_solves_per_s'],1))"; }
run 8; run 12; run 16
timeout 120 python tools/asm_only.py 5 40 3 2>&1 | tail -1
