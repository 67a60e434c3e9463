#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/gpu_gemm_cmp.py 8192 256 8192 1 2>&1 | tail -2 | cut -c1-200
for P in 0 444 888; do echo "== MFB_GEMM_C_PREFETCH=$P"; MFB_GEMM_C_PREFETCH=$P timeout 100 python tools/gpu_gemm.py 20480 256 2>&1 | tail -2 | head -1; MFB_GEMM_C_PREFETCH=$P timeout 100 python tools/gpu_gemm.py 8192 256 2>&1 | tail -2 | head -1; done
for P in 0 444; do echo "== LU MFB_GEMM_C_PREFETCH=$P"; MFB_GEMM_C_PREFETCH=$P timeout 300 python tools/gpu_lu.py time 40 2>&1 | tail -2 | cut -c1-330; done
