#!/bin/bash
mkdir -p gpurun_out
for s in "8192 256 8192" "16402 256 16434" "4096 256 30002"; do timeout 280 python tools/gpu_gemm_cmp.py $s 3 2>&1 | tail -12; done > gpurun_out/r02_gemm_cmp_after_fix.log 2>&1; cat gpurun_out/r02_gemm_cmp_after_fix.log | cut -c1-300
for t in 1 0; do MFB_GEMM_TMA=$t timeout 300 python tools/gpu_lu.py time 40 > gpurun_out/r02_lu_time_tma$t.v2.log 2>&1; cat gpurun_out/r02_lu_time_tma$t.v2.log; done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_v3.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu_v3.log; tail -4 gpurun_out/r02_pytest_gpu_v3.log
