#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/asm_only.py 5 40 4 2>&1 | tail -2
for R in 64 256 512; do echo "== MFB_LU_CLUSTER_MIN_ROWS=$R"; MFB_LU_CLUSTER_MIN_ROWS=$R timeout 120 python tools/c1_step.py 2>&1 | tail -1 | cut -c1-420
  MFB_LU_CLUSTER_MIN_ROWS=$R timeout 300 python bench.py --workload c1 --steps 2 --warmup 3 --lanes 8 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   lanes', d.get('lanes'), 'one at a time', d['one_at_a_time']['e2e_solves_per_s'])"
done
for L in 2 4; do MFB_LU_CLUSTER_MIN_ROWS=256 timeout 300 python bench.py --workload c1 --steps 2 --warmup 3 --lanes $L --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   lanes', d.get('lanes'))"; done
for G in 148 64; do echo "== MFB_LU_PANEL_CTAS=$G"; MFB_LU_PANEL_CTAS=$G timeout 300 python tools/gpu_lu.py time 40 2>&1 | tail -2 | head -1 | cut -c1-330; done
