#!/bin/bash
# NOT RUN in round 2 (GPU budget spent before these rows existed): first GPU validation of the widened rows.
#   gpurun --timeout 900 -- 'bash tools/gpu_r2o.sh'
# 1. the widened GPU tests, then the whole -m gpu suite; 2. the post-pass leg of bench.py alone and the default bench line; 3. ncu: launch list of
# the post leg + one --set full capture of k_taa and k_path_trace (summaries -> profiles/ with tools/ncu_summary.py).
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/widened -m gpu -q 2>&1 | tail -25 > gpurun_out/r2o_tests_widened.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2o_tests_all.log
timeout 300 python bench.py --post-leg 3840 2160 262144 > gpurun_out/r2o_post_leg.json 2> gpurun_out/r2o_post_leg.err
timeout 900 python bench.py > gpurun_out/r2o_bench3.json 2> gpurun_out/r2o_bench3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2o_launches.csv python bench.py --post-leg 3840 2160 262144 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_taa|k_path_trace|k_tonemap' -c 6 -o gpurun_out/r2o_post_full python bench.py --post-leg 3840 2160 262144 > /dev/null 2>&1
ncu -i gpurun_out/r2o_post_full.ncu-rep --page raw --csv > gpurun_out/r2o_post_full_raw.csv 2>/dev/null
