#!/bin/bash
mkdir -p gpurun_out
# the refill of the TMA kernel changed (non-blocking): full comparison first, under a short timeout; on any failure the rest runs with the cp.async kernel
timeout 120 python tools/gpu_gemm_cmp.py 8192 256 8192 2 > gpurun_out/r02_gemm_cmp_v3.log 2>&1; rc=$?
timeout 120 python tools/gpu_gemm_cmp.py 4096 256 30002 2 >> gpurun_out/r02_gemm_cmp_v3.log 2>&1; rc2=$?
cat gpurun_out/r02_gemm_cmp_v3.log | cut -c1-200
if [ $rc -ne 0 ] || [ $rc2 -ne 0 ] || grep -q "bad entries [1-9]" gpurun_out/r02_gemm_cmp_v3.log; then echo "TMA kernel check FAILED (rc $rc $rc2): falling back to MFB_GEMM_TMA=0"; export MFB_GEMM_TMA=0; fi
timeout 60 python tools/gpu_gemm.py 20480 256 2>&1 | tail -2
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_v5.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu_v5.log; tail -25 gpurun_out/r02_pytest_gpu_v5.log | cut -c1-300
timeout 120 python tools/one_step.py 9 20 > gpurun_out/r02_one_step_quad9.log 2>&1; cat gpurun_out/r02_one_step_quad9.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_regular_bulk --launch-skip 3 -c 1 -o gpurun_out/r02_ncu_k1_quad9 -f python tools/asm_only.py 9 20 3 > gpurun_out/r02_ncu_k1_quad9.log 2>&1; tail -2 gpurun_out/r02_ncu_k1_quad9.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pot_regular --launch-skip 1 -c 1 -o gpurun_out/r02_ncu_pot_quad9 -f python tools/acoustic_step.py 9 20 2 > gpurun_out/r02_ncu_pot_quad9.log 2>&1; tail -2 gpurun_out/r02_ncu_pot_quad9.log
