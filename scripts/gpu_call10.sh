#!/bin/bash
# round 2, GPU call 10: ComParE_2016 functionals end to end + the whole GPU suite
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_functionals_gpu.py -m gpu -q -s -k compare16 2>&1 | tail -30 | cut -c1-3000 | tee gpurun_out/c10_compare_func.txt
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/c10_gpu_suite.txt
