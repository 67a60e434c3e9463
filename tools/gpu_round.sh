#!/bin/bash
# One GPU call: parity tests, bench, ncu launch list, ncu --set full captures of the top kernels.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/one_step.py 5 40 > gpurun_out/launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_regular_bulk -c 2 -o gpurun_out/prof_regular -f python tools/one_step.py 5 40 > gpurun_out/prof_regular.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_zgemm3m -s 1 -c 1 -o gpurun_out/prof_zgemm -f python tools/gpu_gemm.py 8192 256 > gpurun_out/prof_zgemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_subpanel|k_trsm_lu' -s 40 -c 2 -o gpurun_out/prof_panel -f python tools/one_step.py 5 30 > gpurun_out/prof_panel.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json
