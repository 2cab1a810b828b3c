#!/bin/bash
# round 2, GPU call 14 (8 GPUs): e2e scaling of the cfg-2 bench with NUMA-bound ranks
set -x
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -14 > gpurun_out/c14_topo.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --no-others --steps 10 --warmup 3 2> gpurun_out/c14_bench8.err | tail -1 > gpurun_out/c14_bench8.json
tail -c 1200 gpurun_out/c14_bench8.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/c14_bench8.json").read())
print("N=8 value %.1f M  e2e %.1f M  ms %.4f" % (l["value"]/1e6, l["e2e"]["value"]/1e6, l["ms_per_step"]), l["e2e"].get("pcie_gbs_per_rank"), l["config"].get("numa"))
PY
