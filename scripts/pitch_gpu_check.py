"""GPU diagnostic: the shipped ComParE_2016.conf (full LLD set, 130 columns) through the session layer against
the reference goldens; prints per-column deviations (run under gpurun)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opensmile_b200.session import Session  # noqa: E402
from opensmile_b200.synth import mixed_pcm, voiced_pcm  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "pitch_goldens.npz"))
conf = os.path.join(ROOT, "oracle", "_ref", "config", "compare16", "ComParE_2016.conf")
cases = {"v32k": voiced_pcm(32000, 16000, seed=7), "m48k": mixed_pcm(48000, 16000, seed=2), "m30k": mixed_pcm(30000, 16000, seed=4),
         "m64k": mixed_pcm(64000, 16000, seed=3), "short_960": voiced_pcm(960, 16000, seed=7), "short_1600": voiced_pcm(1600, 16000, seed=7),
         "short_2400": voiced_pcm(2400, 16000, seed=7)}
s = Session(conf, options={"lldcsvoutput": "x.csv"}, device=0)
names = s.element_names(16000.0, 1)
np.set_printoptions(linewidth=220, precision=6, suppress=True)
# all cases in ONE batch (ragged utterances)
keys = list(cases)
pcm = np.concatenate([cases[k] for k in keys])
off = np.cumsum([0] + [cases[k].size for k in keys]).astype(np.int64)
rows, fo = s.extract_pcm(pcm, off, 16000.0, 1)
for i, k in enumerate(keys):
    got, ref = rows[fo[i]:fo[i + 1]], G[k + "_lld"]
    print("==", k, got.shape, ref.shape)
    if got.shape != ref.shape:
        continue
    sc = np.abs(ref).max(axis=0) + 1e-30
    err = np.abs(got - ref) / sc
    worst = err.max(axis=0)
    for c in np.argsort(-worst)[:8]:
        bad = np.argwhere(err[:, c] > 1e-5)[:, 0]
        print("  %-40s max rel %.3g  rows>1e-5: %d %s" % (names[c], worst[c], bad.size, bad[:10]))
    print("  columns over 1e-5:", int((worst > 1e-5).sum()), "of", worst.size)
