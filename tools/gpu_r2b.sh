#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gi_refl.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2b_tests.log
timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2b_bench3.json 2> gpurun_out/r2b_bench3.err
HR_REFL_ATROUS_IMPL=1 timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2b_bench3_dense.json 2>> gpurun_out/r2b_bench3.err
