#!/bin/bash
# round 2, GPU call 35: whole GPU suite at HEAD (melspec scales, ARFF device sink, NArelTh / Samples / DCT, batched formant steps)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|Error|assert|FAILED" | cut -c1-400 | tail -30 | tee gpurun_out/c35_gpu_suite.txt
