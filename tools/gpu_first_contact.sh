#!/bin/bash
# First hardware contact of the device code that was written without GPU access: every never-run test module goes in
# its OWN python process under its own timeout, so that a device fault in one of them neither poisons the CUDA context of the others nor
# hangs the box; the validated suite runs first and is the regression check of the neutral api.cu edits.  Logs land in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/gpu_first_contact.sh'
mkdir -p gpurun_out
run() { local name=$1; shift; timeout 300 "$@" > gpurun_out/fc_$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/fc_summary.log; }
: > gpurun_out/fc_summary.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/fc_validated.log 2>&1; echo "validated exit $?" | tee -a gpurun_out/fc_summary.log
run interior_pressures python -m pytest tests/test_gpu_acoustic.py -q -m gpu -k interior_pressures
run static_internal    python -m pytest tests/test_gpu_driver.py -q -m gpu -k with_internal_points
run coupled            python -m pytest tests/test_gpu_coupled.py -q -m gpu
run poro_tri3          compute-sanitizer --error-exitcode 9 python -m pytest tests/test_gpu_poroelastic.py -q -m gpu -x -k "0.3-5-3"
run poroelastic        python -m pytest tests/test_gpu_poroelastic.py -q -m gpu
run golden_widening   python -m pytest tests/test_golden_widening.py -q -m gpu
run acoustic_bench     python bench.py --workload acoustic --steps 3 --warmup 3
run coupled_bench      python bench.py --workload coupled --steps 2 --warmup 3
tail -n 3 gpurun_out/fc_*.log | tail -n 60
cat gpurun_out/fc_summary.log
