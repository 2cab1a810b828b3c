#!/bin/bash
# round 2, GPU call 30: whole GPU suite at HEAD (Onset / Peaks / Crossings functionals, sample formats, summary file route)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|Error|assert|FAILED" | cut -c1-400 | tail -30 | tee gpurun_out/c30_gpu_suite.txt
