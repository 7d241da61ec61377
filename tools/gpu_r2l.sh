#!/bin/bash
# final 1-GPU validation of the round: full GPU suite, default bench line (all legs), config 2, two A/Bs, ncu launch list + full capture
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2l_tests.log
timeout 900 python bench.py > gpurun_out/r2l_bench3.json 2> gpurun_out/r2l_bench3.err
timeout 300 python bench.py --config 2 --no-cpu-baseline > gpurun_out/r2l_bench2.json 2> gpurun_out/r2l_bench2.err
HR_FORCE_PEER_TEMPORAL=1 timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 20 > gpurun_out/r2l_bench3_peer_k14.json 2> gpurun_out/r2l_ab.err
HR_REFL_TRACE_MINB=16 timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 20 > gpurun_out/r2l_bench3_minb16.json 2>> gpurun_out/r2l_ab.err
HR_REFL_TRACE_MINB=12 timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 20 > gpurun_out/r2l_bench3_minb12.json 2>> gpurun_out/r2l_ab.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2l_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2l_ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_refl --launch-skip 180 --launch-count 6 -f -o gpurun_out/r2l_full python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2l_ncu_full.log 2>&1
