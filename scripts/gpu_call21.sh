#!/bin/bash
# round 2, GPU call 21: compute-sanitizer over the kernels of this round (lld512_kernel racecheck + memcheck, functionals / sinks memcheck), functionals timing
set -x
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "golden or ragged or tile" 2>&1 | tail -12 | cut -c1-400 | tee gpurun_out/c21_lld512_racecheck.txt
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "golden or ragged or tile or short" 2>&1 | tail -8 | cut -c1-400 | tee gpurun_out/c21_lld512_memcheck.txt
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_functionals_gpu.py tests/test_sinks_gpu.py -m gpu -q -x 2>&1 | tail -8 | cut -c1-400 | tee gpurun_out/c21_functionals_sinks_memcheck.txt
timeout 600 python scripts/time_functionals.py 2000 2>&1 | tail -3 | tee gpurun_out/c21_time_functionals.txt
