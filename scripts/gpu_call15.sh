#!/bin/bash
# round 2, GPU call 15: pitch chain (jitter: independent partial sums; shs: forward scan without recomputation): tests, timing, ncu
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pitch_gpu.py tests/test_zzz_gemaps_gpu.py tests/test_functionals_gpu.py -m gpu -q -x 2>&1 | tail -6 | cut -c1-2000 | tee gpurun_out/c15_pitch_tests.txt
timeout 900 python bench.py --workload compare16 --no-others --steps 3 --warmup 2 2> gpurun_out/c15_bench_compare16.err | tail -1 > gpurun_out/c15_bench_compare16.json
python - <<'PY'
import json
l=json.loads(open("gpurun_out/c15_bench_compare16.json").read())
print("compare16 value %.2f M ms %.1f parity %s" % (l["value"]/1e6, l["ms_per_step"], l.get("parity")))
print(l["roofline"]["kernels_ms"])
PY
for k in jitter_kernel shs_kernel; do
OSM_BENCH_N_UTT=500 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o gpurun_out/c15_$k python bench.py --workload compare16 --no-others --steps 1 --warmup 1 > gpurun_out/c15_${k}_ncu.log 2>&1
done
ls -la gpurun_out | tail -4
