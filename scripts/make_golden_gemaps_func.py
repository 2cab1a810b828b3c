"""Generate tests/golden/gemaps_func.npz with the UNMODIFIED reference (oracle/_ref/SMILExtract): the functionals rows (-csvoutput) of
the shipped config/egemaps/v02/eGeMAPSv02.conf (88 features) and config/gemaps/v01b/GeMAPSv01b.conf (62 features) for three inputs (mixed_pcm(24000, seed 3), voiced_pcm(32000, seed 7),
the reference's recording opensmile.wav resampled to 16 kHz).  Build container only."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refrun  # noqa: E402
from opensmile_b200.synth import mixed_pcm, voiced_pcm  # noqa: E402
from make_golden_functionals import csv_rows  # noqa: E402

REF = "/root/reference/config"


def main():
    rec = np.load(os.path.join(ROOT, "tests", "golden", "egemaps_recordings.npz"))
    rng = np.random.RandomState(11)
    # degenerate contours: digital silence, unvoiced noise only, an utterance shorter than the Viterbi buffer, a single voiced burst
    burst = np.zeros(20000, np.int16)
    burst[6000:12000] = voiced_pcm(6000, 16000, seed=5)
    sigs = {"m24k": mixed_pcm(24000, 16000, seed=3), "v32k": voiced_pcm(32000, 16000, seed=7), "rec": rec["pcm_opensmile_16k"],
            "silence": np.zeros(16000, np.int16), "noise": (rng.randn(16000) * 800).astype(np.int16), "short": voiced_pcm(4000, 16000, seed=9),
            "burst": burst}
    out = {}
    for tag, conf in (("egemaps", "egemaps/v02/eGeMAPSv02.conf"), ("gemaps", "gemaps/v01b/GeMAPSv01b.conf")):
        for key, pcm in sigs.items():
            with tempfile.TemporaryDirectory() as d:
                wav = os.path.join(d, "in.wav")
                refrun.write_wav(wav, pcm, 16000, 1)
                subprocess.run([refrun.SMILEXTRACT, "-C", os.path.join(REF, conf), "-I", wav, "-csvoutput", os.path.join(d, "f.csv"),
                                "-l", "0"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                n, r = csv_rows(os.path.join(d, "f.csv"))
                if key == "m24k":
                    out["csv_" + tag + "_m24k"] = np.frombuffer(open(os.path.join(d, "f.csv"), "rb").read(), np.uint8)   # the sink's file as written
                out["names_" + tag] = np.array(n)
                out[tag + "_" + key] = r
    for k, v in out.items():
        print(k, v.shape)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "gemaps_func.npz"), **out)


if __name__ == "__main__":
    main()
