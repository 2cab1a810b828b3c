#!/bin/bash
# round 2, GPU call 13: whole GPU suite (harmonics lag memo, functionals e2e), then the default bench line (all four workloads)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -12 | cut -c1-3000 | tee gpurun_out/c13_gpu_suite.txt
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/c13_bench.json 2> gpurun_out/c13_bench.err; echo "bench exit $?"
tail -c 1500 gpurun_out/c13_bench.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/c13_bench.json").read().strip().splitlines()[-1])
print("mfcc12 value %.1f M e2e %.1f M cpu %.2f M parity %s" % (l["value"]/1e6, l["e2e"]["value"]/1e6, l["cpu_baseline"]["value"]/1e6, l["parity"]["ok"]))
for o in l.get("other_workloads", []):
    print(o["config"]["workload"][:40], "value %.2f M e2e %.2f M ms %.1f" % (o["value"]/1e6, o["e2e"]["value"]/1e6, o["ms_per_step"]), o["parity"]["ok"], o["roofline"]["kernels_ms"], o.get("cpu_baseline",{}).get("value"))
PY
