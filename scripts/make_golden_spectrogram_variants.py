"""Generate tests/golden/spectrogram_variants.npz with the UNMODIFIED reference (oracle/_ref/SMILExtract): the cFFTmagphase level in
its output variants (tests/configs/spectrogram_variants.conf) for one 0.5 s signal.  Build container only."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refrun  # noqa: E402
from opensmile_b200.synth import mixed_pcm  # noqa: E402

VARIANTS = {"mag": [], "specdens": ["-normalise", "1"], "powspec": ["-power", "1"], "powspecdens": ["-normalise", "1", "-power", "1"],
            "dbpsd": ["-dB", "1"], "dbpsd_floor": ["-dB", "1", "-dBpnorm", "60.0", "-mindBp", "-20.0"]}


def main():
    pcm = mixed_pcm(8000, 16000, seed=5)
    pcm[3000:3400] = 0                                     # digital silence inside: log10(0) behind the dB floor
    conf = os.path.join(ROOT, "tests", "configs", "spectrogram_variants.conf")
    out = {"pcm": pcm}
    with tempfile.TemporaryDirectory() as d:
        wav = os.path.join(d, "in.wav")
        refrun.write_wav(wav, pcm, 16000, 1)
        for name, opts in VARIANTS.items():
            o, c = os.path.join(d, name + ".htk"), os.path.join(d, name + ".csv")
            r = subprocess.run([refrun.SMILEXTRACT, "-C", conf, "-I", wav, "-O", o, "-csvoutput", c, "-l", "1"] + opts, capture_output=True, text=True)
            if r.returncode:
                print(r.stderr[-2000:])
                sys.exit(1)
            rows, _ = refrun.read_htk(o)
            out["rows_" + name] = rows.astype(np.float32)
            out["name0_" + name] = np.array(open(c).readline().strip().split(";")[2])
            print(name, rows.shape, float(rows.min()), float(rows.max()), out["name0_" + name])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "spectrogram_variants.npz"), **out)


if __name__ == "__main__":
    main()
