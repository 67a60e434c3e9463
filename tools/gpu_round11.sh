#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_c1.py tests/test_gpu_parity.py -m gpu -x -q -k "c1 or lanes or sweep or parity or golden or nan" > gpurun_out/r02_pytest_graph.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_graph.log; tail -6 gpurun_out/r02_pytest_graph.log | cut -c1-300
timeout 120 python tools/c1_step.py 2>&1 | tail -1 | cut -c1-300
for L in 1 4 8 16; do timeout 300 python bench.py --workload c1 --steps 2 --warmup 3 --lanes $L --no-cpu-baseline 2>gpurun_out/r02_c1_lanes$L.err | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lanes $L', d.get('lanes'), 'one at a time', d.get('one_at_a_time', d['e2e']['value']), 'value', d['value'])"; done
MFB_LU_GRAPH_MAX_N=0 timeout 300 python bench.py --workload c1 --steps 2 --warmup 3 --lanes 8 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('no graph: lanes 8', d.get('lanes'), 'one at a time', d.get('one_at_a_time'))"
