"""A/B timing of the cfg-2 fused kernel across library variants (opensmile_b200/variants/lib_*.so built with -D switches) and the
default library: one subprocess per library (OSM_B200_LIB), 1 M frames device-resident, CUDA events over 20 launches after 5 warm-ups."""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch
from opensmile_b200 import Plan, components_mfcc12_0_d_a
n_utt, L = 2000, 80240
plan = Plan(components_mfcc12_0_d_a(16000.0), "lld", 0)
g = torch.Generator(device="cuda").manual_seed(0)
pcm = (torch.randn(n_utt * L, device="cuda", generator=g) * 3000).clamp(-32768, 32767).to(torch.int16)
off = np.arange(n_utt + 1, dtype=np.int64) * L
out = plan.run_device(pcm, off)
for _ in range(5):
    plan.run_device(pcm, off, d_out=out)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    plan.run_device(pcm, off, d_out=out)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
print("%%.4f ms  %%.1f M frames/s  checksum %%.6e" %% (ms, out.shape[0] / ms / 1e3, float(out.double().abs().sum())))
''' % ROOT
libs = [("default", None)] + [(os.path.basename(p), p) for p in sorted(glob.glob(os.path.join(ROOT, "opensmile_b200", "variants", "lib_*.so")))]
for name, path in libs:
    env = dict(os.environ)
    if path:
        env["OSM_B200_LIB"] = path
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print("%-28s %s" % (name, (r.stdout.strip().splitlines() or [r.stderr[-300:]])[-1]))
