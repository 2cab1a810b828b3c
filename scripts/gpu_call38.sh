#!/bin/bash
# round 2, GPU call 38: very short utterances against the reference; ComParE end-to-end arm with jitter_kernel at 5 / 6 CTAs per SM
mkdir -p gpurun_out
timeout 600 python scripts/dbg_short_utts.py 2>&1 | tail -22 | cut -c1-400 | tee gpurun_out/c38_short_utts.txt
: > gpurun_out/c38_e2e_ab.txt
for v in default lib_jit5.so default; do
  if [ "$v" = default ]; then unset OSM_B200_LIB; else export OSM_B200_LIB=$PWD/opensmile_b200/variants/$v; fi
  timeout 600 python bench.py --workload compare16 --no-others --steps 5 --warmup 3 2> gpurun_out/c38_$v.err | tail -1 > gpurun_out/c38_$v.json
  python - "$v" <<'PY' | tee -a gpurun_out/c38_e2e_ab.txt
import json, sys
l = json.loads(open("gpurun_out/c38_%s.json" % sys.argv[1]).read())
print("%-14s value %.2f M rows/s (%.1f ms)  e2e %.2f M  pcie %s" % (sys.argv[1], l["value"] / 1e6, l["ms_per_step"], l["e2e"]["value"] / 1e6, l["e2e"]["pcie_gbs_per_rank"]))
PY
done
unset OSM_B200_LIB
