#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1; echo "ncu exit $?"
python tools/summarize_ncu.py gpurun_out/r02_launches.csv > gpurun_out/r02_launches.md 2>&1; head -30 gpurun_out/r02_launches.md
rm -f gpurun_out/r02_launches.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_regular_bulk --launch-skip 3 -c 1 -o gpurun_out/r02_ncu_k1_tri3 -f python tools/asm_only.py 5 40 3 > gpurun_out/r02_ncu_k1_tri3.log 2>&1; tail -1 gpurun_out/r02_ncu_k1_tri3.log
