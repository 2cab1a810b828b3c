#!/bin/bash
# round 2, GPU call 41: ComParE end-to-end arm with the float-sample read path compiled out of the kernels (A/B, same box)
mkdir -p gpurun_out
: > gpurun_out/c41_nof32_ab.txt
for v in lib_nof32.so default; do
  if [ "$v" = default ]; then unset OSM_B200_LIB; else export OSM_B200_LIB=$PWD/opensmile_b200/variants/$v; fi
  timeout 600 python bench.py --workload compare16 --no-others --steps 5 --warmup 3 2> /dev/null | tail -1 > gpurun_out/c41_$v.json
  python - "$v" <<'PY' | tee -a gpurun_out/c41_nof32_ab.txt
import json, sys
l = json.loads(open("gpurun_out/c41_%s.json" % sys.argv[1]).read())
k = l["roofline"]["kernels_ms"]
print("%-14s value %.2f M rows/s (%.1f ms)  e2e %.2f M  kernels sum %.1f  energy %.2f mzcr %.2f jitter %.2f" % (sys.argv[1], l["value"] / 1e6, l["ms_per_step"], l["e2e"]["value"] / 1e6, sum(k.values()), k["energy_kernel"], k["mzcr_kernel"], k["jitter_kernel"]))
PY
done
