#!/bin/bash
# round 2, GPU call 29: sample-format tests with full assertion output
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pcm_formats.py -m gpu -q 2>&1 | grep -E "passed|failed|AssertionError|Error|assert" | cut -c1-300 | tee gpurun_out/c29_pcm_formats.txt
