#!/bin/bash
# round 2, GPU call 26: full GPU suite + smoke at HEAD (lld512 unroll defaults, GeMAPS summaries), default bench, lld512 ncu capture, summary timing
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | cut -c1-2500 | tee gpurun_out/c26_gpu_suite.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee gpurun_out/c26_smoke.txt
timeout 1200 python bench.py --steps 10 --warmup 3 2> gpurun_out/c26_bench.err | tail -1 > gpurun_out/c26_bench.json
tail -c 1500 gpurun_out/c26_bench.json
timeout 600 python scripts/time_functionals.py 2000 egemaps 2>&1 | tail -3 | tee gpurun_out/c26_time_egemaps_func.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lld512 -c 1 -o gpurun_out/c26_lld512 python bench.py --no-others --steps 1 --warmup 1 > gpurun_out/c26_lld512_ncu.log 2>&1
ls -la gpurun_out/c26_lld512.ncu-rep
