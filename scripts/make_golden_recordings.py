"""Generate tests/golden/egemaps_recordings.npz with the UNMODIFIED reference (oracle/_ref/SMILExtract) on the reference's own
recordings (example-audio/opensmile.wav, media-interpretation.wav; 44.1 kHz mono):

    python scripts/make_golden_recordings.py      # build container only (needs /root/reference and `make -C oracle ref`)

For each recording two inputs: the first 2.5 s at the native 44.1 kHz, and the whole recording resampled to 16 kHz
(scipy.signal.resample_poly 160/441, rounded to int16) -- the rate the GeMAPS sets are specified for.  Stored per input:
pcm_<key> (int16), sr_<key>, egemaps_<key> = rows of config/egemaps/v02/eGeMAPSv02.conf -lldhtkoutput (25 columns),
compare_<key> = rows of config/compare16/ComParE_2016.conf -lldhtkoutput joined with its lld_de level (130 columns, 16 kHz
inputs only)."""
import os
import subprocess
import sys
import tempfile

import numpy as np
from scipy.signal import resample_poly

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refrun  # noqa: E402


def ref_rows(conf, pcm, sr, opt):
    with tempfile.TemporaryDirectory() as d:
        wav = os.path.join(d, "in.wav")
        refrun.write_wav(wav, pcm, sr, 1)
        out = os.path.join(d, "l.htk")
        subprocess.run([refrun.SMILEXTRACT, "-C", os.path.join(refrun.CONFIG_DIR, conf), "-I", wav, opt, out, "-l", "0"],
                       check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        return refrun.read_htk(out)[0]


def main():
    assert refrun.available()
    out = {}
    for name in ("opensmile", "media-interpretation"):
        pcm, sr, nch = refrun.read_wav(os.path.join("/root/reference/example-audio", name + ".wav"))
        assert nch == 1 and sr == 44100
        key = name.replace("-", "_")
        x16 = np.clip(np.round(resample_poly(pcm.astype(np.float64), 160, 441)), -32768, 32767).astype(np.int16)
        for k, x, r in ((key + "_16k", x16, 16000), (key + "_44k1", pcm[:110250].copy(), 44100)):
            out["pcm_" + k] = x
            out["sr_" + k] = np.int64(r)
            out["egemaps_" + k] = ref_rows("egemaps/v02/eGeMAPSv02.conf", x, r, "-lldhtkoutput")
            if r == 16000:
                out["compare_" + k] = ref_rows("compare16/ComParE_2016.conf", x, r, "-lldhtkoutput")
            print(k, x.shape, out["egemaps_" + k].shape, out.get("compare_" + k, np.zeros(0)).shape)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "egemaps_recordings.npz"), **out)


if __name__ == "__main__":
    main()
