#!/bin/bash
# 8-GPU box: NCCL tests, scaling lines N=8/4/2 with the overlapped gather (point-to-point), N=8 with the broadcast gather for A/B
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2k_gpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k nccl 2>&1 | tail -25 > gpurun_out/r2k_tests.log
run() { # n port tag extra-args...
  n=$1; port=$2; tag=$3; shift 3
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r2k_bench_$tag.json 2> gpurun_out/r2k_bench_$tag.err
  echo rc=$?
}
run 8 29521 n8
HR_GATHER_IMPL=0 run 8 29522 n8_bcast --no-extras
run 4 29523 n4 --no-extras
run 2 29524 n2 --no-extras
