#!/bin/bash
# GPU call for the widened paths: tests, static bench, ncu launch list + captures of the static kernels
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_static.csv python tools/static_step.py 9 11 2 > gpurun_out/launches_static.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_regular_bulk' -s 3 -c 2 -o gpurun_out/prof_static_k1 -f python tools/static_step.py 5 40 2 > gpurun_out/prof_static_k1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_dgemm_minus' -s 300 -c 1 -o gpurun_out/prof_dgemm -f python tools/static_step.py 5 40 1 > gpurun_out/prof_dgemm.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
