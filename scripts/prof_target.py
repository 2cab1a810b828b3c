"""Short driver for ncu captures: the BASELINE cfg-2 batch through the device entry point."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from opensmile_b200 import Plan, components_mfcc12_0_d_a
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
plan = Plan(components_mfcc12_0_d_a(16000.0), "lld", 0)
pcm = bench.synth_batch_torch(bench.N_UTT, bench.UTT_LEN, torch.device("cuda", 0), 0)
off = np.arange(bench.N_UTT + 1, dtype=np.int64) * bench.UTT_LEN
out = None
for i in range(n):
    out = plan.run_device(pcm, off, d_out=out)
torch.cuda.synchronize()
print("done", out.shape, plan.last_kernel_times())
