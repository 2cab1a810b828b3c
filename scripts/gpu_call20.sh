#!/bin/bash
# round 2, GPU call 20: formant inverse sum with 128-bit loads, spectral_kernel with 8 warps, new windows (description only): suite + eGeMAPS / ComParE timing
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 | cut -c1-2500 | tee gpurun_out/c20_gpu_suite.txt
for w in egemaps compare16; do
timeout 900 python bench.py --workload $w --no-others --steps 3 --warmup 2 2> gpurun_out/c20_bench_$w.err | tail -1 > gpurun_out/c20_bench_$w.json
python - <<PY
import json
l=json.loads(open("gpurun_out/c20_bench_$w.json").read())
print("$w value %.2f M ms %.1f parity %s" % (l["value"]/1e6, l["ms_per_step"], l.get("parity",{}).get("ok")))
print(l["roofline"]["kernels_ms"])
PY
done
