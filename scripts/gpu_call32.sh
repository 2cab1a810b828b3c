#!/bin/bash
# round 2, GPU call 32: compute-sanitizer memcheck over the kernels added late in the round (pcm_convert_kernel, gated seq_post_kernel,
# summary_assemble_kernel, Onset / Peaks / Crossings in functionals_kernel), then the whole GPU suite at HEAD
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_pcm_formats.py -m gpu -q -x 2>&1 | tail -6 | cut -c1-300 | tee gpurun_out/c32_pcm_formats_memcheck.txt
timeout 1500 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_functionals_gpu.py -m gpu -q -x -k "gemaps or onset or degenerate or summary" 2>&1 | tail -6 | cut -c1-300 | tee gpurun_out/c32_summaries_memcheck.txt
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|Error|assert|FAILED" | cut -c1-400 | tail -20 | tee gpurun_out/c32_gpu_suite.txt
