"""Generate tests/golden/formant_goldens.npz with the UNMODIFIED reference (oracle/_ref/SMILExtract):

    python scripts/make_golden_formant.py        # needs `make -C oracle ref` (build container only)

Level taps of tests/configs/formant_taps.conf (the GeMAPS formant chain) on mixed_pcm(24000, seed=3):
  res [T, 220]  cSpecResample output (11 kHz frames)      lpc [T, 11]  cLpc coefficients
  fmt [T, 10]   cFormantLpc: formantFreqLpc[1..5] | formantBandwidthLpc[1..5]
and of tests/configs/harmonics_taps.conf (same input): h_f0 [T60, 3] Viterbi level (F0final first), h_fmt, h_mag [T60, 513]
60 ms magnitude spectrum, h_harm [T60, 6] cHarmonics: HNRdBACF, H1-H2, H1-A3, F1..F3 amplitude (log rel. F0)
and of tests/configs/gemaps_vq_taps.conf (the shipped GeMAPSv01b_core.lld.conf.inc unchanged): g_f0 = gemapsv01b_logPitch
[T60, 3], g_jit = gemapsv01b_jitterShimmer [T60, 2], g_fmt = gemapsv01b_formants [T25, 10], g_harm = gemapsv01b_harmonics [T60, 6]
and the LLD file of the shipped config/gemaps/v01b/GeMAPSv01b.conf (-lldhtkoutput, 18 columns): gemaps_lld_m24k (same input),
gemaps_lld_m40k = mixed_pcm(40000, seed=5); the same two inputs through config/egemaps/v02/eGeMAPSv02.conf (25 columns):
egemaps_lld_m24k, egemaps_lld_m40k, and the column names of that file (names_egemaps_lld, from its -lldcsvoutput header)
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refrun  # noqa: E402
from opensmile_b200.synth import mixed_pcm  # noqa: E402


def main():
    assert refrun.available(), "build the reference first: make -C oracle ref"
    pcm = mixed_pcm(24000, 16000, seed=3)
    out = {}
    with tempfile.TemporaryDirectory() as d:
        refrun.write_wav(os.path.join(d, "in.wav"), pcm, 16000, 1)
        subprocess.run([refrun.SMILEXTRACT, "-C", os.path.join(ROOT, "tests", "configs", "formant_taps.conf"), "-I", "in.wav", "-l", "0"],
                       cwd=d, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        for k in ("res", "lpc", "fmt"):
            out[k] = refrun.read_htk(os.path.join(d, k + ".htk"))[0]
    with tempfile.TemporaryDirectory() as d:            # cHarmonics with its three input levels
        refrun.write_wav(os.path.join(d, "in.wav"), pcm, 16000, 1)
        subprocess.run([refrun.SMILEXTRACT, "-C", os.path.join(ROOT, "tests", "configs", "harmonics_taps.conf"), "-I", "in.wav", "-l", "0"],
                       cwd=d, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        for k in ("f0", "mag", "harm"):
            out["h_" + k] = refrun.read_htk(os.path.join(d, k + ".htk"))[0]
        out["h_fmt"] = refrun.read_htk(os.path.join(d, "fmt.htk"))[0]
    taps = open(os.path.join(ROOT, "tests", "configs", "gemaps_vq_taps.conf")).read().replace("REFCONF", refrun.CONFIG_DIR)
    with tempfile.TemporaryDirectory() as d:            # the shipped GeMAPS graph itself, its four voice-quality levels
        refrun.write_wav(os.path.join(d, "in.wav"), pcm, 16000, 1)
        open(os.path.join(d, "t.conf"), "w").write(taps)
        subprocess.run([refrun.SMILEXTRACT, "-C", "t.conf", "-I", "in.wav", "-l", "0"], cwd=d, check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        for k in ("f0", "jit", "fmt", "harm"):
            out["g_" + k] = refrun.read_htk(os.path.join(d, k + ".htk"))[0]
    full = os.path.join(refrun.CONFIG_DIR, "gemaps", "v01b", "GeMAPSv01b.conf")      # the shipped feature set, LLD sink
    for name, x in (("gemaps_lld_m24k", pcm), ("gemaps_lld_m40k", mixed_pcm(40000, 16000, seed=5))):
        with tempfile.TemporaryDirectory() as d:
            wav = os.path.join(d, "in.wav")
            refrun.write_wav(wav, x, 16000, 1)
            subprocess.run([refrun.SMILEXTRACT, "-C", full, "-I", wav, "-lldhtkoutput", os.path.join(d, "l.htk"), "-l", "0"],
                           check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            out[name] = refrun.read_htk(os.path.join(d, "l.htk"))[0]
    efull = os.path.join(refrun.CONFIG_DIR, "egemaps", "v02", "eGeMAPSv02.conf")
    for name, x in (("egemaps_lld_m24k", pcm), ("egemaps_lld_m40k", mixed_pcm(40000, 16000, seed=5))):
        with tempfile.TemporaryDirectory() as d:
            wav = os.path.join(d, "in.wav")
            refrun.write_wav(wav, x, 16000, 1)
            subprocess.run([refrun.SMILEXTRACT, "-C", efull, "-I", wav, "-lldhtkoutput", os.path.join(d, "l.htk"),
                            "-lldcsvoutput", os.path.join(d, "l.csv"), "-l", "0"],
                           check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            out[name] = refrun.read_htk(os.path.join(d, "l.htk"))[0]
            out["names_egemaps_lld"] = np.array(open(os.path.join(d, "l.csv")).readline().strip().split(";")[2:])
    print({k: v.shape for k, v in out.items()})
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "formant_goldens.npz"), **out)


if __name__ == "__main__":
    main()
