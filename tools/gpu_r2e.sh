#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2e_tests.log
timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2e_bench3.json 2> gpurun_out/r2e_bench3.err
HR_REFL_ATROUS_IMPL=2 timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2e_bench3_impl2.json 2>> gpurun_out/r2e_bench3.err
