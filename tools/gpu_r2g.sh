#!/bin/bash
# 2-GPU call: reflections multi-rank tests + NCCL tests + scaling bench lines (with torchrun diagnostics)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "reflections or nccl" 2>&1 | tail -25 > gpurun_out/r2g_tests.log
export PYTHONFAULTHANDLER=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 --log-dir gpurun_out/r2g_torchrun --tee 3 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err
echo "rc=$?" >> gpurun_out/r2g_bench_n2.err
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err
timeout 600 python -m pytest tests/test_gpu_config_sizes.py tests/test_gbuffer.py tests/test_gpu_headless.py tests/test_gpu_spp.py tests/test_gpu_gi_refl.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2g_tests2.log
