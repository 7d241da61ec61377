#!/bin/bash
# after the K12 register retune + push / K14 scheduling changes: emulated-rank tests, MINB 18 / 20 A/B, final default line
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_gi_refl.py tests/test_gpu_config_sizes.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2m_tests.log
for m in 18 20; do
  HR_REFL_TRACE_MINB=$m timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 20 > gpurun_out/r2m_bench3_minb$m.json 2>> gpurun_out/r2m_ab.err
done
timeout 900 python bench.py > gpurun_out/r2m_bench3.json 2> gpurun_out/r2m_bench3.err
