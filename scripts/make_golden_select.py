"""Generate tests/golden/select_goldens.npz with the UNMODIFIED reference (oracle/_ref/SMILExtract):

    python scripts/make_golden_select.py        # needs `make -C oracle ref` (build container only)

tests/configs/gemaps_sel.conf (cDataSelector on top of the shipped GeMAPS graph) on mixed_pcm(24000, seed=3) and
voiced_pcm(32000, seed=7): gsel_m24k / gsel_v32k [T, 8] rows of its CSV file, gsel_names = the header's element names.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refrun  # noqa: E402
from opensmile_b200.synth import mixed_pcm, voiced_pcm  # noqa: E402


def main():
    assert refrun.available(), "build the reference first: make -C oracle ref"
    conf = open(os.path.join(ROOT, "tests", "configs", "gemaps_sel.conf")).read().replace("REFCONF", refrun.CONFIG_DIR)
    out = {}
    for key, pcm in (("gsel_m24k", mixed_pcm(24000, 16000, seed=3)), ("gsel_v32k", voiced_pcm(32000, 16000, seed=7))):
        with tempfile.TemporaryDirectory() as d:
            refrun.write_wav(os.path.join(d, "in.wav"), pcm, 16000, 1)
            open(os.path.join(d, "t.conf"), "w").write(conf)
            subprocess.run([refrun.SMILEXTRACT, "-C", "t.conf", "-I", "in.wav", "-O", "o.csv", "-l", "0"], cwd=d, check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            lines = open(os.path.join(d, "o.csv")).read().splitlines()
        out["gsel_names"] = np.array(lines[0].split(";")[2:])
        out[key] = np.array([ln.split(";")[2:] for ln in lines[1:]], np.float64).astype(np.float32)
        out[key + "_time"] = np.array([ln.split(";")[1] for ln in lines[1:]], np.float64)
    print({k: v.shape for k, v in out.items()}, list(out["gsel_names"]))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "select_goldens.npz"), **out)


if __name__ == "__main__":
    main()
