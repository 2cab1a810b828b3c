#!/bin/bash
# First device run of the formant kernel (opensmile_b200/csrc/formant.cu); use under gpurun:
#   gpurun --timeout 900 -- 'bash scripts/formant_gpu_check.sh'
# 1. the gated GPU tests, 2. the same under compute-sanitizer memcheck, 3. kernel time of one 10 000-utterance batch.
set -x
mkdir -p gpurun_out
export OSM_B200_RUN_UNVERIFIED=1
timeout 600 python -m pytest tests/test_zz_select_gpu.py tests/test_zzz_formant_gpu.py tests/test_zzz_gemaps_gpu.py -q 2>&1 | tee gpurun_out/formant_gpu_tests.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_zzz_formant_gpu.py -x -q -k host_build \
  > gpurun_out/formant_memcheck.txt 2>&1; echo "memcheck exit $?" >> gpurun_out/formant_memcheck.txt
tail -5 gpurun_out/formant_memcheck.txt
timeout 600 python scripts/bench_general.py tests/configs/formant_chain.conf 2000 48000 \
  2>&1 | tee gpurun_out/formant_bench.txt
