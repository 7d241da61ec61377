#!/bin/bash
# A/B of the traversal variants (compile-time: Makefile VARIANT=, selected with HR_BUILD_DIR)
set -x
mkdir -p gpurun_out
HR_BUILD_DIR=build_w1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gbuffer.py tests/test_gpu_gi_refl.py tests/test_gpu_trace_pt.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2i_tests_w1.log
for v in build_t0 build build_w1 build_l2 build_w1l2 build_w1l1; do
  HR_BUILD_DIR=$v timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r2i_c3_$v.json 2> gpurun_out/r2i_c3_$v.err
done
for v in build_t0 build build_w1 build_w1l2; do
  HR_BUILD_DIR=$v timeout 300 python bench.py --config 2 --no-extras --no-cpu-baseline --steps 30 --warmup 5 > gpurun_out/r2i_c2_$v.json 2> gpurun_out/r2i_c2_$v.err
done
