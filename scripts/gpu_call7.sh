#!/bin/bash
# round 2, GPU call 7: lld512_kernel v4 -- reference-order DCT (default) vs fused partial-sum DCT (variant library)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_session_gpu.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/c7_gpu_suite.txt
for v in default fused_dct; do
  if [ $v = default ]; then unset OSM_B200_LIB; else export OSM_B200_LIB=$PWD/opensmile_b200/variants/lib_$v.so; fi
  timeout 600 python bench.py --no-others --steps 20 --warmup 3 2> gpurun_out/c7_bench_$v.err | tail -1 > gpurun_out/c7_bench_$v.json
  python - <<PY
import json
l=json.loads(open("gpurun_out/c7_bench_$v.json").read())
print("$v value %.1f M  ms %.4f  e2e %.1f M parity %s" % (l["value"]/1e6, l["ms_per_step"], l["e2e"]["value"]/1e6, l.get("parity")))
PY
done
unset OSM_B200_LIB
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lld512 -c 1 -o gpurun_out/c7_lld512 python bench.py --no-others --steps 1 --warmup 1 > gpurun_out/c7_ncu.log 2>&1
ls -la gpurun_out | tail -3
