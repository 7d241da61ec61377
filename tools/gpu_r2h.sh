#!/bin/bash
# 8-GPU call: one-process-per-GPU tests at world 8, scaling bench lines at N = 8 and N = 4 (on the same box)
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2h_gpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "nccl" 2>&1 | tail -25 > gpurun_out/r2h_tests.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2h_bench_n8.json 2> gpurun_out/r2h_bench_n8.err
echo "rc=$?" >> gpurun_out/r2h_bench_n8.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2h_bench_n4.json 2> gpurun_out/r2h_bench_n4.err
echo "rc=$?" >> gpurun_out/r2h_bench_n4.err
