#!/bin/bash
# round 2, GPU call A: full GPU test suite, config-3 / config-2 bench lines, ncu launch list + one full capture of a config-3 frame
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r2a_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2a_tests.log
timeout 600 python bench.py > gpurun_out/r2a_bench3.json 2> gpurun_out/r2a_bench3.err
timeout 300 python bench.py --config 2 --no-cpu-baseline > gpurun_out/r2a_bench2.json 2> gpurun_out/r2a_bench2.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2a_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_refl --launch-skip 180 --launch-count 6 -f -o gpurun_out/r2a_full python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2a_ncu_full.log 2>&1
ls -la gpurun_out
