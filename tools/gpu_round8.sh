#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_v5.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu_v5.log; tail -25 gpurun_out/r02_pytest_gpu_v5.log | cut -c1-300
timeout 120 python tools/one_step.py 9 20 > gpurun_out/r02_one_step_quad9.log 2>&1; cat gpurun_out/r02_one_step_quad9.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_regular_bulk --launch-skip 3 -c 1 -o gpurun_out/r02_ncu_k1_quad9 -f python tools/asm_only.py 9 20 3 > gpurun_out/r02_ncu_k1_quad9.log 2>&1; tail -2 gpurun_out/r02_ncu_k1_quad9.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pot_regular --launch-skip 1 -c 1 -o gpurun_out/r02_ncu_pot_quad9 -f python tools/acoustic_step.py 9 20 2 > gpurun_out/r02_ncu_pot_quad9.log 2>&1; tail -2 gpurun_out/r02_ncu_pot_quad9.log
