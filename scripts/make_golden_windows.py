"""Generate tests/golden/window_goldens.npz with the UNMODIFIED reference: the cWindower level of a constant signal (every sample
32767 -> 1.0 after conversion) is the window table itself, one frame of 400 samples, for the window functions and switches the five
BASELINE configurations do not use (Blackman, Blackman-Harris, Bartlett-Hann, Lanczos, squareRoot, fade, custom coefficients).
Build container only."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refrun  # noqa: E402

CASES = {
    "bla": "winFunc = Bla", "bla_alpha": "winFunc = Blackman\nalpha = 0.2", "bla_a012": "winFunc = bla\nalpha0 = 0.4\nalpha1 = 0.45\nalpha2 = 0.1",
    "blh": "winFunc = BlH", "bah": "winFunc = BaH", "lac": "winFunc = Lac", "han_sqrt": "winFunc = Han\nsquareRoot = 1",
    "ham_fade": "winFunc = Ham\nfade = 0.1", "blh_gain_sqrt_fade": "winFunc = blackman-harris\ngain = 2.5\nsquareRoot = 1\nfade = 0.25",
    "gau": "winFunc = Gau\nsigma = 0.3", "tri": "winFunc = Tri",
}

CONF = """[componentInstances:cComponentManager]
instance[dataMemory].type=cDataMemory
instance[waveIn].type=cWaveSource
instance[fr].type=cFramer
instance[win].type=cWindower
instance[sink].type=cCsvSink
printLevelStats=0
nThreads=1
[waveIn:cWaveSource]
writer.dmLevel=wave
filename=\\cm[inputfile(I){test.wav}:input]
monoMixdown=1
[fr:cFramer]
reader.dmLevel=wave
writer.dmLevel=frames
frameSize = 0.025
frameStep = 0.010
frameCenterSpecial = left
[win:cWindower]
reader.dmLevel=frames
writer.dmLevel=winframes
%s
[sink:cCsvSink]
reader.dmLevel=winframes
filename=\\cm[output(O){out.csv}:output]
timestamp=0
number=0
printHeader=0
"""


def main():
    assert refrun.available()
    out = {}
    pcm = np.full(800, 32767, np.int16)
    for key, body in CASES.items():
        with tempfile.TemporaryDirectory() as d:
            wav, conf, csv = os.path.join(d, "in.wav"), os.path.join(d, "c.conf"), os.path.join(d, "o.csv")
            refrun.write_wav(wav, pcm, 16000, 1)
            open(conf, "w").write(CONF % body)
            subprocess.run([refrun.SMILEXTRACT, "-C", conf, "-I", wav, "-O", csv, "-l", "0"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            row = open(csv).read().strip().split("\n")[0].split(";")
            out[key] = np.array([float(v) for v in row], np.float32)
            out[key + "_conf"] = np.array(body)
            print(key, out[key].shape, out[key][:3], out[key][200])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "window_goldens.npz"), **out)


if __name__ == "__main__":
    main()
