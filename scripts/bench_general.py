"""Device-resident throughput of a general (multi-kernel) plan built from a .conf file.
usage: bench_general.py [conf] [n_utt] [n_samples] [sample_rate] [n_channels]
(default: tests/configs/compare_ns.conf, 1000 utterances x 48000 sample frames, 16 kHz mono; the synthetic
signal is generated for 16 kHz -- other rates only change the frame geometry, multi-channel input is the
mono signal reshaped, which is fine for timing)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from opensmile_b200 import Plan, Session
import bench

conf = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "configs", "compare_ns.conf")
n_utt = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
n_samp = int(sys.argv[3]) if len(sys.argv) > 3 else 48000
sr = float(sys.argv[4]) if len(sys.argv) > 4 else 16000.0
nch = int(sys.argv[5]) if len(sys.argv) > 5 else 1
# OSM_BENCH_OPTS="lldcsvoutput=x.csv,other=value": command line options of the configuration file (e.g. the shipped
# ComParE_2016.conf only has an LLD sink when -lldcsvoutput is given)
opts = dict(kv.split("=", 1) for kv in os.environ.get("OSM_BENCH_OPTS", "").split(",") if "=" in kv)
s = Session(conf, options=opts or None, device=-1)
comps, level = s.components(sr, nch)
plan = Plan(list(comps), level, 0)
pcm = bench.synth_batch_torch(n_utt, n_samp * nch, torch.device("cuda", 0), 0)
off = np.arange(n_utt + 1, dtype=np.int64) * n_samp
out = plan.run_device(pcm, off)
torch.cuda.synchronize()
for _ in range(3):
    out = plan.run_device(pcm, off, d_out=out)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
K = 10
for _ in range(K):
    out = plan.run_device(pcm, off, d_out=out)
ev1.record()
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / K
rows = out.shape[0]
print("%s @%g Hz x%d: %d rows x %d cols, %.3f ms/step, %.1f M rows/s, launches %d" % (os.path.basename(conf), sr, nch, rows, out.shape[1], ms, rows / ms / 1e3, plan.last_launch_count()))
