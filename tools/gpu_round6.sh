#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zgemm3m_tma --launch-skip 1 -c 1 -o gpurun_out/r02_ncu_zgemm_tma_20k -f python tools/gpu_gemm.py 20480 256 > gpurun_out/r02_ncu_zgemm_tma_20k.log 2>&1; tail -2 gpurun_out/r02_ncu_zgemm_tma_20k.log
for nb in 256 384 512; do MFB_LU_NB=$nb timeout 300 python tools/gpu_lu.py time 40 2>&1 | tail -2 | sed "s/^/nb=$nb /"; done > gpurun_out/r02_lu_nb_sweep.log 2>&1; cat gpurun_out/r02_lu_nb_sweep.log
