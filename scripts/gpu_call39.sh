#!/bin/bash
# round 2, GPU call 39: ComParE end-to-end arm, the tree of call 26 (commit 2675a20, copied to gpurun_old/) against HEAD on the same box
mkdir -p gpurun_out
: > gpurun_out/c39_e2e_old_vs_head.txt
for t in old head old head; do
  if [ "$t" = old ]; then d=gpurun_old; else d=.; fi
  (cd $d && timeout 600 python bench.py --workload compare16 --no-others --steps 5 --warmup 3 2> /dev/null | tail -1) > gpurun_out/c39_$t.json
  python - "$t" <<'PY' | tee -a gpurun_out/c39_e2e_old_vs_head.txt
import json, sys
l = json.loads(open("gpurun_out/c39_%s.json" % sys.argv[1]).read())
print("%-5s value %.2f M rows/s (%.1f ms)  e2e %.2f M  pcie %s" % (sys.argv[1], l["value"] / 1e6, l["ms_per_step"], l["e2e"]["value"] / 1e6, l["e2e"]["pcie_gbs_per_rank"]))
PY
done
