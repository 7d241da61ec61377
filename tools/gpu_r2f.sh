#!/bin/bash
# 2-GPU call: multi-GPU tests (NCCL, one process per GPU) + scaling bench lines
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2f_gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_config_sizes.py tests/test_gbuffer.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2f_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err
