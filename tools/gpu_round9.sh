#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_v6.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu_v6.log; tail -25 gpurun_out/r02_pytest_gpu_v6.log | cut -c1-300
timeout 120 python tools/asm_only.py 5 40 4 > gpurun_out/r02_asm_only_tri3_gridconst.log 2>&1; tail -3 gpurun_out/r02_asm_only_tri3_gridconst.log
for L in 8 16; do timeout 300 python bench.py --workload c1 --steps 2 --warmup 3 --lanes $L --no-cpu-baseline > gpurun_out/r02_bench_c1_lanes$L.json 2> gpurun_out/r02_bench_c1_lanes$L.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_c1_lanes$L.json').read().strip().splitlines()[-1])
print('lanes $L', d.get('lanes'), 'one at a time', d.get('one_at_a_time'))
PY
done
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_headline_v1.json 2> gpurun_out/r02_bench_headline_v1.err; tail -c 600 gpurun_out/r02_bench_headline_v1.json
