#!/bin/bash
# round 2, GPU call 40: bisecting the ComParE end-to-end regression (trees of two intermediate commits copied beside HEAD)
mkdir -p gpurun_out
: > gpurun_out/c40_bisect.txt
for t in 7f15a1c 35c2f0f head; do
  if [ "$t" = head ]; then d=.; else d=gpurun_$t; fi
  (cd $d && OSM_BENCH_SKIP_CPU=1 timeout 600 python bench.py --workload compare16 --no-others --steps 5 --warmup 3 2> /dev/null | tail -1) > gpurun_out/c40_$t.json
  python - "$t" <<'PY' | tee -a gpurun_out/c40_bisect.txt
import json, sys
l = json.loads(open("gpurun_out/c40_%s.json" % sys.argv[1]).read())
print("%-8s value %.2f M rows/s (%.1f ms)  e2e %.2f M" % (sys.argv[1], l["value"] / 1e6, l["ms_per_step"], l["e2e"]["value"] / 1e6))
PY
done
