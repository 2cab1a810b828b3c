#!/bin/bash
# round 2, GPU call 9: cFunctionals Times / Lpc / Segments / Peaks2 on the device; formant kernel with two frames per warp
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_functionals_gpu.py -m gpu -q 2>&1 | tail -40 | tee gpurun_out/c9_functionals.txt
timeout 900 python -m pytest tests/test_zzz_formant_gpu.py tests/test_zzz_gemaps_gpu.py tests/test_zz_select_gpu.py -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/c9_formant.txt
timeout 900 python bench.py --workload egemaps --no-others --steps 3 --warmup 2 2> gpurun_out/c9_bench_egemaps.err | tail -1 > gpurun_out/c9_bench_egemaps.json
python - <<'PY'
import json
l=json.loads(open("gpurun_out/c9_bench_egemaps.json").read())
print("egemaps value %.2f M ms %.1f parity %s" % (l["value"]/1e6, l["ms_per_step"], l.get("parity")))
print(l["roofline"]["kernels_ms"])
PY
