#!/bin/bash
# round 2, GPU call 3: lane-parallel harmonics kernel, the reworked bench (all four BASELINE workloads, parity inside, NUMA binding)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/c3_gpu_suite.txt
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err; echo "bench exit $?"
tail -c 3000 gpurun_out/c3_bench.err
python - <<'PY'
import json
try:
    l=json.loads(open("gpurun_out/c3_bench.json").read().strip().splitlines()[-1])
    print("mfcc12 value %.1f M e2e %.1f M cpu %.2f M (per-process %.2f M) parity %s" % (l["value"]/1e6, l["e2e"]["value"]/1e6, l["cpu_baseline"]["value"]/1e6, l["cpu_baseline"].get("per_process_value",0)/1e6, l["parity"]))
    for o in l.get("other_workloads", []):
        print(o["config"]["workload"][:40], "value %.2f M e2e %.2f M ms %.1f" % (o["value"]/1e6, o["e2e"]["value"]/1e6, o["ms_per_step"]), o["parity"], o["roofline"]["kernels_ms"], o.get("cpu_baseline",{}).get("value"))
except Exception as e:
    print("parse failed", e)
PY
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/c3_bench_reference.json | cut -c1-900
OSM_BENCH_N_UTT=500 timeout 900 ncu --set full --clock-control none --import-source on -k regex:harmonics_kernel -c 1 -o gpurun_out/c3_harmonics_kernel \
    python bench.py --workload egemaps --no-others --steps 1 --warmup 3 > gpurun_out/c3_harmonics_kernel_ncu.log 2>&1
ls -la gpurun_out | tail -8
