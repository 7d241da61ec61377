#!/bin/bash
# config 2 regression hunt: shadow packet traversal on/off, new default build (TRI_MODE 0, LEAF_MAX 2)
set -x
mkdir -p gpurun_out
for pk in 1 0; do
  HR_SHADOW_PACKET=$pk timeout 300 python bench.py --config 2 --no-extras --no-cpu-baseline --steps 30 --warmup 5 > gpurun_out/r2j_c2_packet$pk.json 2> gpurun_out/r2j_c2_packet$pk.err
done
HR_SHADOW_PACKET=0 HR_TRACE_IMPL=1 timeout 300 python bench.py --config 2 --no-extras --no-cpu-baseline --steps 30 --warmup 5 > gpurun_out/r2j_c2_pt.json 2> gpurun_out/r2j_c2_pt.err
timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r2j_c3.json 2> gpurun_out/r2j_c3.err
timeout 300 python bench.py --config 1 --no-extras --no-cpu-baseline --steps 30 --warmup 5 > gpurun_out/r2j_c1.json 2> gpurun_out/r2j_c1.err
