#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2d_tests.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2d_bench3.json 2> gpurun_out/r2d_bench3.err
HR_REFL_ATROUS_MINB=3 timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2d_bench3_minb3.json 2>> gpurun_out/r2d_bench3.err
HR_REFL_ATROUS_MINB=2 timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2d_bench3_minb2.json 2>> gpurun_out/r2d_bench3.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_refl --launch-skip 180 --launch-count 6 -f -o gpurun_out/r2d_full python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2d_ncu_full.log 2>&1
