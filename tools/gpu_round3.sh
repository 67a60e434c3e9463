#!/bin/bash
# round 2, GPU batch 3: first run of the TMA trailing-update kernel (correctness on ragged shapes, then speed against the cp.async kernel), LU at 30k, C1
mkdir -p gpurun_out
L=gpurun_out/r02_gemm_tma_first.log; : > $L
for shape in "1000 48 777" "64 16 32" "130 32 70" "2048 256 2048"; do
  timeout 120 python tools/gpu_gemm.py $shape >> $L 2>&1 || echo "FAILED shape $shape rc=$?" >> $L
done
for t in 1 0; do
  MFB_GEMM_TMA=$t timeout 200 python tools/gpu_gemm.py 8192 256 >> $L 2>&1
  MFB_GEMM_TMA=$t timeout 300 python tools/gpu_gemm.py 20480 256 >> $L 2>&1
done
cat $L
for t in 1 0; do MFB_GEMM_TMA=$t timeout 300 python tools/gpu_lu.py time 40 > gpurun_out/r02_lu_time_tma$t.log 2>&1; cat gpurun_out/r02_lu_time_tma$t.log; done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_v2.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu_v2.log; tail -4 gpurun_out/r02_pytest_gpu_v2.log
timeout 300 python tools/c1_step.py > gpurun_out/r02_c1_step_v2.log 2>&1; cat gpurun_out/r02_c1_step_v2.log
timeout 600 python bench.py --workload c1 --steps 2 --warmup 3 > gpurun_out/r02_bench_c1_v1.json 2> gpurun_out/r02_bench_c1_v1.err; tail -c 1500 gpurun_out/r02_bench_c1_v1.json
