#!/bin/bash
# round 2, GPU call 22: A/B timing of lld512_kernel unroll variants
set -x
mkdir -p gpurun_out
timeout 900 python scripts/ab_lld512.py 2>&1 | tee gpurun_out/c22_ab.txt
