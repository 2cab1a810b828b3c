#!/bin/bash
# round 2, GPU call 11: device sinks, ComParE_2016 functionals end to end (padded union rows, Viterbi-dependent frame counts), whole suite
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sinks_gpu.py -m gpu -q 2>&1 | tail -15 | cut -c1-2000 | tee gpurun_out/c11_sinks.txt
OSM_B200_DEBUG_FUNC=1 timeout 900 python -m pytest tests/test_functionals_gpu.py -m gpu -q -s -k compare16 2>&1 | tail -30 | cut -c1-6000 | tee gpurun_out/c11_compare_func.txt
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/c11_gpu_suite.txt
