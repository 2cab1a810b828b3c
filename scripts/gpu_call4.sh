#!/bin/bash
# round 2, GPU call 4 (re-entry): whole GPU suite incl. the cFunctionals tests, smoke, default bench
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/c4_gpu_suite.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/c4_smoke.txt
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err; echo "bench exit $?"
tail -c 2000 gpurun_out/c4_bench.err
cut -c1-1500 gpurun_out/c4_bench.json
