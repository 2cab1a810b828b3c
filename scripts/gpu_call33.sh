#!/bin/bash
# round 2, GPU call 33: Samples / DCT on the GPU, smoke() with the summary check, default bench line with the summaries key
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_functionals_gpu.py -m gpu -q 2>&1 | grep -E "passed|failed|Error|assert|FAILED" | cut -c1-400 | tail -12 | tee gpurun_out/c33_functionals.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -5 | tee gpurun_out/c33_smoke.txt
timeout 1500 python bench.py --steps 10 --warmup 3 2> gpurun_out/c33_bench.err | tail -1 > gpurun_out/c33_bench.json
python - <<'PY'
import json
l = json.loads(open("gpurun_out/c33_bench.json").read())
print("value %.1f M  e2e %.1f M  frac %.4f  parity %s" % (l["value"] / 1e6, l["e2e"]["value"] / 1e6, l["roofline"]["frac"], l["parity"]["ok"]))
for o in l.get("other_workloads", []): print(o["config"]["workload"][:40], "%.2f M" % (o["value"] / 1e6), o["parity"]["ok"])
print(json.dumps(l.get("summaries")))
PY
