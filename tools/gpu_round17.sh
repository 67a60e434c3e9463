#!/bin/bash
mkdir -p gpurun_out
MFB_GEMM_TMA_CFG=1 timeout 120 python tools/gpu_gemm_cmp.py 8192 256 8192 1 2>&1 | tail -2 | cut -c1-200
for C in 0 1; do echo "== MFB_GEMM_TMA_CFG=$C"; MFB_GEMM_TMA_CFG=$C timeout 100 python tools/gpu_gemm.py 20480 256 2>&1 | tail -2 | head -1; MFB_GEMM_TMA_CFG=$C timeout 100 python tools/gpu_gemm.py 8192 256 2>&1 | tail -2 | head -1; done
echo "== LU MFB_GEMM_TMA_CFG=1"; MFB_GEMM_TMA_CFG=1 timeout 300 python tools/gpu_lu.py time 40 2>&1 | tail -2 | cut -c1-330
