#!/bin/bash
# round 2, GPU call 6: lld512_kernel v3 (chunk context in smem, DCT partial sums fused into the mel phase)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_session_gpu.py -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/c6_gpu_suite.txt
for fast in 1; do
  OSM_B200_LLD_FAST=$fast timeout 600 python bench.py --no-others --steps 20 --warmup 3 2> gpurun_out/c6_bench_fast$fast.err | tail -1 > gpurun_out/c6_bench_fast$fast.json
  python - <<PY
import json
l=json.loads(open("gpurun_out/c6_bench_fast$fast.json").read())
print("fast=$fast value %.1f M  ms %.4f  e2e %.1f M parity %s" % (l["value"]/1e6, l["ms_per_step"], l["e2e"]["value"]/1e6, l.get("parity")))
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lld512 -c 1 -o gpurun_out/c6_lld512 python bench.py --no-others --steps 1 --warmup 1 > gpurun_out/c6_ncu.log 2>&1
ls -la gpurun_out | tail -3
