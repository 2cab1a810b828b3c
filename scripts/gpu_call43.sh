#!/bin/bash
# round 2, GPU call 43 (last): whole GPU suite and the default bench line at HEAD
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|Error|assert|FAILED" | cut -c1-300 | tail -10 | tee gpurun_out/c43_gpu_suite.txt
timeout 900 python bench.py --steps 10 --warmup 3 2> gpurun_out/c43_bench.err | tail -1 > gpurun_out/c43_bench.json
python - <<'PY'
import json
l = json.loads(open("gpurun_out/c43_bench.json").read())
print("value %.1f M  e2e %.1f M  frac %.4f  parity %s" % (l["value"] / 1e6, l["e2e"]["value"] / 1e6, l["roofline"]["frac"], l["parity"]["ok"]))
for o in l.get("other_workloads", []): print(o["config"]["workload"][:40], "%.2f M  e2e %.2f M" % (o["value"] / 1e6, o["e2e"]["value"] / 1e6), o["parity"]["ok"])
for s in l.get("summaries", []): print(s.get("config"), s.get("utterances_per_s"), s.get("error"))
PY
