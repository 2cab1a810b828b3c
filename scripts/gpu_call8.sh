#!/bin/bash
# round 2, GPU call 8: cFunctionals Times / Lpc / Segments / Peaks2 on the device
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_functionals_gpu.py -m gpu -q 2>&1 | tail -40 | tee gpurun_out/c8_functionals.txt
