#!/bin/bash
# round 2, GPU call 27: jitter_kernel occupancy A/B (min blocks per SM 5 / 6 / 8) on the ComParE workload
mkdir -p gpurun_out
./opensmile_b200/variants/probe_f32x2 | tee gpurun_out/c27_probe_f32x2.txt
: > gpurun_out/c27_jitter_ab.txt
for v in default lib_jit5.so lib_jit6.so lib_jit8.so; do
  if [ "$v" = default ]; then unset OSM_B200_LIB; else export OSM_B200_LIB=$PWD/opensmile_b200/variants/$v; fi
  timeout 600 python bench.py --workload compare16 --no-others --steps 3 --warmup 2 2> gpurun_out/c27_$v.err | tail -1 > gpurun_out/c27_$v.json
  python - "$v" <<'PY' | tee -a gpurun_out/c27_jitter_ab.txt
import json, sys
v = sys.argv[1]
l = json.loads(open("gpurun_out/c27_%s.json" % v).read())
k = l["roofline"]["kernels_ms"]
print("%-14s value %.2f M rows/s  ms %.1f  jitter %.2f shs %.2f viterbi %.2f parity %s" % (v, l["value"] / 1e6, l["ms_per_step"], k.get("jitter_kernel", -1), k.get("shs_kernel", -1), k.get("viterbi_kernel", -1), l.get("parity", {}).get("ok")))
PY
done
unset OSM_B200_LIB
OSM_B200_FUNC_TIMING=1 timeout 600 python scripts/time_functionals.py 2000 egemaps 2>&1 | tail -6 | tee gpurun_out/c27_time_egemaps_func.txt
OSM_B200_FUNC_TIMING=1 timeout 600 python scripts/time_functionals.py 2000 compare16 2>&1 | tail -6 | tee gpurun_out/c27_time_compare16_func.txt
