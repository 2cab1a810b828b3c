#!/bin/bash
# round 2, GPU call 19: ncu captures of formant_kernel / harmonics_kernel / spectral_kernel at HEAD
set -x
mkdir -p gpurun_out
for k in formant_kernel harmonics_kernel spectral_kernel; do
OSM_BENCH_N_UTT=1000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o gpurun_out/c19_$k python bench.py --workload egemaps --no-others --steps 1 --warmup 1 > gpurun_out/c19_${k}_ncu.log 2>&1
done
ls -la gpurun_out | tail -4
