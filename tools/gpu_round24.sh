#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_size" 2>&1 | tail -3
timeout 900 python bench.py --steps 3 --warmup 3 --two-seam --no-cpu-baseline > gpurun_out/r02_bench_two_seam.json 2> gpurun_out/r02_bench_two_seam.err; echo "exit $?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_two_seam.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], d.get('two_seam'))
PY
