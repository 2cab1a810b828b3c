#!/bin/bash
# round 2, GPU call 17: ncu captures of the pitch-chain kernels at HEAD (source-line attribution)
set -x
mkdir -p gpurun_out
for k in jitter_kernel shs_kernel; do
OSM_BENCH_N_UTT=1000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o gpurun_out/c17_$k python bench.py --workload compare16 --no-others --steps 1 --warmup 1 > gpurun_out/c17_${k}_ncu.log 2>&1
done
ls -la gpurun_out | tail -3
