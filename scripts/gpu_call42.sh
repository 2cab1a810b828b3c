#!/bin/bash
# round 2, GPU call 42: two builds of the PCM-reading kernels (int16 / float samples): sample-format + parity + pitch + GeMAPS tests, ComParE line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pcm_formats.py tests/test_parity_gpu.py tests/test_pitch_gpu.py tests/test_zzz_gemaps_gpu.py tests/test_short_utterances_gpu.py -m gpu -q 2>&1 | grep -E "passed|failed|Error|assert|FAILED" | cut -c1-300 | tail -10 | tee gpurun_out/c42_tests.txt
timeout 600 python bench.py --workload compare16 --no-others --steps 5 --warmup 3 2> /dev/null | tail -1 > gpurun_out/c42_compare16.json
python - <<'PY' | tee gpurun_out/c42_compare16.txt
import json
l = json.loads(open("gpurun_out/c42_compare16.json").read())
k = l["roofline"]["kernels_ms"]
print("HEAD value %.2f M rows/s (%.1f ms)  e2e %.2f M  parity %s  energy %.2f mzcr %.2f" % (l["value"] / 1e6, l["ms_per_step"], l["e2e"]["value"] / 1e6, l["parity"]["ok"], k["energy_kernel"], k["mzcr_kernel"]))
PY
