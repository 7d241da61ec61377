#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2c_tests.log
timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2c_bench3.json 2> gpurun_out/r2c_bench3.err
HR_REFL_TRACE_IMPL=0 timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2c_bench3_fused.json 2>> gpurun_out/r2c_bench3.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_refl --launch-skip 210 --launch-count 7 -f -o gpurun_out/r2c_full python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2c_ncu_full.log 2>&1
