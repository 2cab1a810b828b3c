"""Generate tests/golden/conf_goldens.npz: outputs of the UNMODIFIED reference (oracle/_ref/SMILExtract)
for whole configuration files, used by the session (conf front end) tests.

    python scripts/make_golden_conf.py        # needs `make -C oracle ref` (build container only)

Contents (inputs are regenerated in the tests from opensmile_b200.synth.voiced_pcm with the seeds below):
  mix16k            tests/configs/lld_mix.conf, voiced_pcm(16000, 16000, seed=3)           [97, 52]
  mix32k_stereo     tests/configs/lld_mix.conf, voiced_pcm(16000, 32000, seed=4, n_chan=2)   (FFT 1024 + 2048)
  mfcc_e            config/mfcc/MFCC12_E_D_A.conf, voiced_pcm(12000, 16000, seed=5)
  mfcc_e_short_<n>  the same config on the first n samples of seed 5, n = 400, 560, 720, 880 (1..4 frames)
  plp_e             config/plp/PLP_E_D_A.conf, voiced_pcm(12000, 16000, seed=6)
  names_<case>      the CSV header's element names
  htk_bytes / csv_bytes   the reference's HTK and CSV files for case mfcc_e (instance name 'utt7')
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refrun  # noqa: E402
from opensmile_b200.synth import voiced_pcm  # noqa: E402


def run(conf, pcm, sr, nch, keep_files=False):
    with tempfile.TemporaryDirectory() as d:
        wav, htk, csv = (os.path.join(d, x) for x in ("in.wav", "out.htk", "out.csv"))
        refrun.write_wav(wav, pcm, sr, nch)
        subprocess.run([refrun.SMILEXTRACT, "-C", conf, "-I", wav, "-O", htk, "-csvoutput", csv, "-instname", "utt7", "-l", "0"],
                       check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        rows, _ = refrun.read_htk(htk)
        names = open(csv).readline().strip().split(";")[2:]
        files = (open(htk, "rb").read(), open(csv, "rb").read()) if keep_files else None
    return rows, np.array(names), files


def main():
    assert refrun.available(), "build the reference first: make -C oracle ref"
    out = {}
    mix = os.path.join(ROOT, "tests", "configs", "lld_mix.conf")
    out["mix16k"], out["names_mix"], _ = run(mix, voiced_pcm(16000, 16000, seed=3), 16000, 1)
    out["mix32k_stereo"], _, _ = run(mix, voiced_pcm(16000, 32000, seed=4, n_chan=2), 32000, 2)
    mfe = os.path.join(refrun.CONFIG_DIR, "mfcc", "MFCC12_E_D_A.conf")
    pcm = voiced_pcm(12000, 16000, seed=5)
    out["mfcc_e"], out["names_mfcc_e"], files = run(mfe, pcm, 16000, 1, keep_files=True)
    out["htk_bytes"] = np.frombuffer(files[0], dtype=np.uint8)
    out["csv_bytes"] = np.frombuffer(files[1], dtype=np.uint8)
    for n in (400, 560, 720, 880):
        out["mfcc_e_short_%d" % n], _, _ = run(mfe, pcm[:n], 16000, 1)
    out["plp_e"], out["names_plp_e"], _ = run(os.path.join(refrun.CONFIG_DIR, "plp", "PLP_E_D_A.conf"),
                                              voiced_pcm(12000, 16000, seed=6), 16000, 1)
    for k, v in out.items():
        print(k, v.shape)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "conf_goldens.npz"), **out)


if __name__ == "__main__":
    main()
