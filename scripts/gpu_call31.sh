#!/bin/bash
# round 2, GPU call 31: warp-per-utterance Viterbi smoother: pitch / formant / functionals tests (bit-exact goldens), then A/B timing on ComParE + eGeMAPS
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pitch_gpu.py tests/test_formant_gpu.py tests/test_functionals_gpu.py -m gpu -q 2>&1 | grep -E "passed|failed|Error|assert|FAILED" | cut -c1-400 | tail -20 | tee gpurun_out/c31_tests.txt
: > gpurun_out/c31_viterbi_ab.txt
for w in compare16 egemaps; do
for v in default lib_vit_thread.so; do
  if [ "$v" = default ]; then unset OSM_B200_LIB; else export OSM_B200_LIB=$PWD/opensmile_b200/variants/$v; fi
  timeout 600 python bench.py --workload $w --no-others --steps 3 --warmup 2 2> gpurun_out/c31_${w}_$v.err | tail -1 > gpurun_out/c31_${w}_$v.json
  python - "$v" "$w" <<'PY' | tee -a gpurun_out/c31_viterbi_ab.txt
import json, sys
v, w = sys.argv[1], sys.argv[2]
l = json.loads(open("gpurun_out/c31_%s_%s.json" % (w, v)).read())
k = l["roofline"]["kernels_ms"]
print("%-10s %-18s value %.2f M rows/s  ms %.1f  jitter %.2f shs %.2f viterbi %.2f parity %s" % (w, v, l["value"] / 1e6, l["ms_per_step"], k.get("jitter_kernel", -1), k.get("shs_kernel", -1), k.get("viterbi_kernel", -1), l.get("parity", {}).get("ok")))
PY
done
done
unset OSM_B200_LIB
