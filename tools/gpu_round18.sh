#!/bin/bash
# 2 GPUs: the C-ABI sweep over real NCCL, then the bench contract at N = 2 (ours, short) and the reference arm under torchrun (rank 0 only, threads pinned)
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sweep_check.py > gpurun_out/r02_sweep_check_2gpu.log 2>&1; echo "sweep_check exit $?" >> gpurun_out/r02_sweep_check_2gpu.log; grep -E "rank|exit" gpurun_out/r02_sweep_check_2gpu.log | cut -c1-250
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench n2 exit $?"; tail -c 1800 gpurun_out/r02_bench_n2.json | cut -c1-1800
