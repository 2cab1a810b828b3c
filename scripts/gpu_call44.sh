#!/bin/bash
# round 2, GPU call 44: the earlier GeMAPS versions' summaries (new test) -- last seconds of the budget
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_functionals_gpu.py -m gpu -q -k "earlier_gemaps" 2>&1 | grep -E "passed|failed|Error|assert|FAILED" | cut -c1-400 | tail -8 | tee gpurun_out/c44_earlier_gemaps.txt
