#!/bin/bash
# last 8-GPU call of the round: scaling lines with the final kernels (K12 at 18 CTAs/SM, split chunk push, K14 halo-first)
set -x
mkdir -p gpurun_out
run() { n=$1; port=$2; tag=$3; to=$4; shift 4
  timeout $to python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r2n_bench_$tag.json 2> gpurun_out/r2n_bench_$tag.err
  echo rc=$?
}
run 8 29531 n8 150
run 4 29532 n4 100 --no-extras
run 2 29533 n2 100 --no-extras
