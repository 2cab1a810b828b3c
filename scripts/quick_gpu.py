"""Quick GPU sanity + timing (dev helper, not the bench)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
g.smoke()
from opensmile_b200 import Plan, components_mfcc12_0_d_a
n_utt, L = 2000, 80240
plan = Plan(components_mfcc12_0_d_a(16000.0), "lld", 0)
gen = torch.Generator(device="cuda").manual_seed(0)
pcm = (torch.randn(n_utt * L, device="cuda", generator=gen) * 3000).clamp(-32768, 32767).to(torch.int16)
off = np.arange(n_utt + 1, dtype=np.int64) * L
out = plan.run_device(pcm, off)
torch.cuda.synchronize()
print("rows", out.shape, "finite", bool(torch.isfinite(out).all()))
for i in range(5):
    plan.run_device(pcm, off, d_out=out)
    torch.cuda.synchronize()
    ms = plan.last_kernel_ms()
    print("kernel ms %.3f -> %.1f Mframes/s, %.1f GB/s algorithmic" % (ms, out.shape[0] / ms / 1e3, out.shape[0] * 476 / ms / 1e6))
