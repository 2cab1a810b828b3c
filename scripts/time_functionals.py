"""Time a summary configuration (LLD plan + cFunctionals instances) end to end from host PCM: ComParE_2016 (6 373 features per
utterance) or eGeMAPSv02 (88):
python scripts/time_functionals.py [n_utt] [compare16|egemaps]   -- dev helper, prints utterances / s with and without the summary."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opensmile_b200.session import Session  # noqa: E402
from opensmile_b200.synth import mixed_pcm  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
which = sys.argv[2] if len(sys.argv) > 2 else "compare16"
conf = os.path.join(ROOT, "oracle", "_ref", "config", *{"compare16": ("compare16", "ComParE_2016.conf"), "egemaps": ("egemaps", "v02", "eGeMAPSv02.conf")}[which])
base = [mixed_pcm(48000, 16000, seed=s) for s in range(8)]
pcm = np.concatenate([base[i % 8] for i in range(n)])
off = np.arange(n + 1, dtype=np.int64) * 48000
for opts, tag in (({"lldcsvoutput": "x.csv"}, "LLD rows only"), ({"csvoutput": "x.csv"}, "LLD + functionals")):
    s = Session(conf, options=opts, device=0)
    s.extract_pcm(pcm[:48000 * 8], off[:9], 16000.0, 1)      # warm-up
    t0 = time.time()
    rows, fo = s.extract_pcm(pcm, off, 16000.0, 1)
    dt = time.time() - t0
    print("%-20s %d utterances x 3 s: %.3f s  -> %.0f utterances/s, output %s" % (tag, n, dt, n / dt, rows.shape))
    s.close()
