#!/bin/bash
# round 2, GPU call 34: formant_kernel with the sequential per-frame steps batched on one warp: formant / GeMAPS / sinks tests, A/B on eGeMAPS
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zzz_formant_gpu.py tests/test_zzz_gemaps_gpu.py tests/test_sinks_gpu.py tests/test_functionals_gpu.py -m gpu -q 2>&1 | grep -E "passed|failed|Error|assert|FAILED" | cut -c1-400 | tail -12 | tee gpurun_out/c34_tests.txt
: > gpurun_out/c34_formant_ab.txt
for v in default lib_fmt_old.so; do
  if [ "$v" = default ]; then unset OSM_B200_LIB; else export OSM_B200_LIB=$PWD/opensmile_b200/variants/$v; fi
  timeout 600 python bench.py --workload egemaps --no-others --steps 3 --warmup 2 2> gpurun_out/c34_$v.err | tail -1 > gpurun_out/c34_$v.json
  python - "$v" <<'PY' | tee -a gpurun_out/c34_formant_ab.txt
import json, sys
v = sys.argv[1]
l = json.loads(open("gpurun_out/c34_%s.json" % v).read())
k = l["roofline"]["kernels_ms"]
print("%-16s value %.2f M rows/s  ms %.1f  formant %.2f harmonics %.2f shs %.2f parity %s" % (v, l["value"] / 1e6, l["ms_per_step"], k.get("formant_kernel", -1), k.get("harmonics_kernel", -1), k.get("shs_kernel", -1), l.get("parity", {}).get("ok")))
PY
done
unset OSM_B200_LIB
