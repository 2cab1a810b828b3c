#!/bin/bash
# round 2, GPU call 28: new tests (sample formats, GeMAPS summaries on degenerate inputs, summary file route), then the whole suite
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pcm_formats.py tests/test_functionals_gpu.py -m gpu -q -x 2>&1 | tail -15 | cut -c1-1500 | tee gpurun_out/c28_new_tests.txt
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | cut -c1-2500 | tee gpurun_out/c28_gpu_suite.txt
