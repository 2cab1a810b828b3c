"""Generate tests/golden/functionals_goldens.npz with the UNMODIFIED reference (oracle/_ref/SMILExtract):

    python scripts/make_golden_functionals.py      # build container only (needs /root/reference and `make -C oracle ref`)

For three inputs (mixed_pcm(24000, seed 3), voiced_pcm(32000, seed 7), the reference's recording opensmile.wav resampled to
16 kHz): the LLD rows (lld;lld_de, 32 columns, exact float32 from -lldhtkoutput; names from the CSV header) and the functionals row (384 values + names) of the shipped
config/is09-13/IS09_emotion.conf, and the three functionals levels of tests/configs/func_variants.conf (names + rows)."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refrun  # noqa: E402
from opensmile_b200.synth import mixed_pcm, voiced_pcm  # noqa: E402

REF = "/root/reference/config"


def csv_rows(path):
    lines = open(path).read().strip().split("\n")
    names = lines[0].split(";")
    rows = [ln.split(";") for ln in lines[1:]]
    first = 2 if names[1] == "frameTime" else 1
    return names[first:], np.array([[float(v) for v in r[first:]] for r in rows], np.float32)


def main():
    assert refrun.available()
    rec = np.load(os.path.join(ROOT, "tests", "golden", "egemaps_recordings.npz"))
    sigs = {"m24k": mixed_pcm(24000, 16000, seed=3), "v32k": voiced_pcm(32000, 16000, seed=7), "rec": rec["pcm_opensmile_16k"]}
    out = {}
    var = open(os.path.join(ROOT, "tests", "configs", "func_variants.conf")).read().replace("REFCONF", REF)
    for key, pcm in sigs.items():
        with tempfile.TemporaryDirectory() as d:
            wav = os.path.join(d, "in.wav")
            refrun.write_wav(wav, pcm, 16000, 1)
            subprocess.run([refrun.SMILEXTRACT, "-C", os.path.join(REF, "is09-13", "IS09_emotion.conf"), "-I", wav, "-csvoutput", os.path.join(d, "f.csv"),
                            "-lldcsvoutput", os.path.join(d, "l.csv"), "-lldhtkoutput", os.path.join(d, "l.htk"), "-l", "0"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            n, r = csv_rows(os.path.join(d, "f.csv"))
            out["is09_func_names"] = np.array(n)
            out["is09_func_" + key] = r
            n, r = csv_rows(os.path.join(d, "l.csv"))
            out["is09_lld_names"] = np.array(n)
            out["is09_lld_" + key] = refrun.read_htk(os.path.join(d, "l.htk"))[0]        # exact float32 rows
            open(os.path.join(d, "v.conf"), "w").write(var)
            subprocess.run([refrun.SMILEXTRACT, "-C", os.path.join(d, "v.conf"), "-I", wav, "-outA", os.path.join(d, "a.csv"), "-outB", os.path.join(d, "b.csv"),
                            "-outC", os.path.join(d, "c.csv"), "-l", "0"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            for lv in "abc":
                n, r = csv_rows(os.path.join(d, lv + ".csv"))
                out["var%s_names" % lv.upper()] = np.array(n)
                out["var%s_%s" % (lv.upper(), key)] = r
    for k, v in out.items():
        print(k, v.shape)
    if "--second-only" not in sys.argv and "--third-only" not in sys.argv:
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "functionals_goldens.npz"), **out)
    # second set (tests/configs/func_variants2.conf): Times, Lpc, Segments, Peaks2
    out2 = {}
    var2 = open(os.path.join(ROOT, "tests", "configs", "func_variants2.conf")).read().replace("REFCONF", REF)
    for key, pcm in sigs.items():
        with tempfile.TemporaryDirectory() as d:
            wav = os.path.join(d, "in.wav")
            refrun.write_wav(wav, pcm, 16000, 1)
            open(os.path.join(d, "v.conf"), "w").write(var2)
            cmd = [refrun.SMILEXTRACT, "-C", os.path.join(d, "v.conf"), "-I", wav, "-l", "0"]
            for lv in "DEFGH":
                cmd += ["-out" + lv, os.path.join(d, lv + ".csv")]
            subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            for lv in "DEFGH":
                n, r = csv_rows(os.path.join(d, lv + ".csv"))
                out2["var%s_names" % lv] = np.array(n)
                out2["var%s_%s" % (lv, key)] = r
    for k, v in out2.items():
        print(k, v.shape)
    if "--third-only" not in sys.argv:
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "functionals_goldens2.npz"), **out2)
    # third set (tests/configs/func_variants3.conf): Onset, Peaks, Crossings
    out3 = {}
    var3 = open(os.path.join(ROOT, "tests", "configs", "func_variants3.conf")).read().replace("REFCONF", REF)
    for key, pcm in sigs.items():
        with tempfile.TemporaryDirectory() as d:
            wav = os.path.join(d, "in.wav")
            refrun.write_wav(wav, pcm, 16000, 1)
            open(os.path.join(d, "v.conf"), "w").write(var3)
            cmd = [refrun.SMILEXTRACT, "-C", os.path.join(d, "v.conf"), "-I", wav, "-l", "0"]
            for lv in "IJKL":
                cmd += ["-out" + lv, os.path.join(d, lv + ".csv")]
            subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            for lv in "IJKL":
                n, r = csv_rows(os.path.join(d, lv + ".csv"))
                out3["var%s_names" % lv] = np.array(n)
                out3["var%s_%s" % (lv, key)] = r
    for k, v in out3.items():
        print(k, v.shape)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "functionals_goldens3.npz"), **out3)


if __name__ == "__main__":
    main()
