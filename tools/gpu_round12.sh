#!/bin/bash
mkdir -p gpurun_out
run() { timeout 200 python bench.py --workload c1 --steps 2 --warmup 3 --lanes $1 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$2 lanes $1:', round(d['lanes']['solves_per_s'],1), 'one at a time', round(d['one_at_a_time']['e2e_solves_per_s'],1))"; }
CUDA_DEVICE_MAX_CONNECTIONS=32 run 4 "maxconn32"
CUDA_DEVICE_MAX_CONNECTIONS=32 run 8 "maxconn32"
CUDA_DEVICE_MAX_CONNECTIONS=32 MFB_LU_LOOKAHEAD=0 run 8 "maxconn32 nolookahead"
MFB_LU_LOOKAHEAD=0 run 8 "nolookahead"
CUDA_DEVICE_MAX_CONNECTIONS=32 MFB_LU_CLUSTER=8 run 8 "maxconn32 cluster8"
