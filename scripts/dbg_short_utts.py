"""GPU check of very short utterances (1 .. 25 frames of the 60 ms level) on the shipped ComParE_2016 / eGeMAPSv02 LLD graphs against the
reference's rows (tests/golden/short_utterances.npz): row counts and per-column error, one line per length."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opensmile_b200.session import Session  # noqa: E402
from opensmile_b200.synth import voiced_pcm  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "short_utterances.npz"))
REFCONF = os.path.join(ROOT, "oracle", "_ref", "config")
LENS = (900, 1000, 1130, 1290, 1450, 1610, 2000, 3000, 4800)
for conf, tag in (("compare16/ComParE_2016.conf", "c16"), ("egemaps/v02/eGeMAPSv02.conf", "ege")):
    s = Session(os.path.join(REFCONF, conf), options={"lldcsvoutput": "x.csv"}, device=0)
    names = s.element_names()
    pcms = [voiced_pcm(n, 16000, seed=n) for n in LENS]
    off = np.concatenate([[0], np.cumsum(LENS)]).astype(np.int64)
    rows, fo = s.extract_pcm(np.concatenate(pcms), off, 16000.0, 1)
    s.close()
    for u, n in enumerate(LENS):
        ref = G["%s_%d" % (tag, n)]
        got = rows[fo[u]:fo[u + 1]]
        if ref.size == 0 or got.shape != ref.shape:
            print(tag, n, "rows got", got.shape, "ref", ref.shape, "OK" if got.shape[0] == ref.shape[0] else "ROW COUNT DIFFERS")
            continue
        scale = np.abs(G["%s_4800" % tag]).max(axis=0) + 1e-9
        err = np.abs(got - ref) / scale
        bad = [(names[c], int(np.argmax(err[:, c])), float(got[np.argmax(err[:, c]), c]), float(ref[np.argmax(err[:, c]), c])) for c in np.nonzero(err.max(axis=0) > 1e-4)[0]]
        print(tag, n, got.shape, "max err/scale %.2e" % err.max(), "bad columns:", len(bad), bad[:4])
