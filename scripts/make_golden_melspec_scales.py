"""Generate tests/golden/melspec_scales.npz with the UNMODIFIED reference (oracle/_ref/SMILExtract): MFCC 0..12 behind cMelspec on every
frequency scale (tests/configs/mfcc_scales.conf: mel, bark, bark_speex, bark_schroed, semitone, linear, log with two bases) for one
1.5 s signal.  Build container only."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refrun  # noqa: E402
from opensmile_b200.synth import mixed_pcm  # noqa: E402

# name -> extra command line options
VARIANTS = {"mel": ["-scale", "mel"], "bark": ["-scale", "bark"], "bark_speex": ["-scale", "bark_speex"], "bark_schroed": ["-scale", "bark_schroed"],
            "semitone": ["-scale", "semitone", "-firstNote", "55.0", "-lofreq", "60"], "linear": ["-scale", "linear"],
            "log2": ["-scale", "log", "-lofreq", "50"], "log10": ["-scale", "log", "-logScaleBase", "10.0", "-lofreq", "50"]}


def main():
    pcm = mixed_pcm(24000, 16000, seed=3)
    conf = os.path.join(ROOT, "tests", "configs", "mfcc_scales.conf")
    out = {}
    with tempfile.TemporaryDirectory() as d:
        wav = os.path.join(d, "in.wav")
        refrun.write_wav(wav, pcm, 16000, 1)
        for name, opts in VARIANTS.items():
            o = os.path.join(d, name + ".htk")
            r = subprocess.run([refrun.SMILEXTRACT, "-C", conf, "-I", wav, "-O", o, "-l", "1"] + opts, capture_output=True, text=True)
            if r.returncode:
                print(r.stderr[-2000:])
                sys.exit(1)
            rows, _ = refrun.read_htk(o)
            out["mfcc_" + name] = rows.astype(np.float32)
            print(name, rows.shape, float(np.abs(rows).max()))
        # the band level itself (cMelspec as the level a sink reads): exact values through an extra cHtkSink, names through a cCsvSink
        for name, opts in (("mel", ["-scale", "mel"]), ("bark", ["-scale", "bark"]), ("htk", ["-melhtk", "1"])):
            txt = open(conf).read() + ("\n[componentInstances:cComponentManager]\ninstance[mh].type=cHtkSink\ninstance[mc].type=cCsvSink\n"
                                       "[mh:cHtkSink]\nreader.dmLevel=melspec\nfilename=%s\nparmKind=9\n"
                                       "[mc:cCsvSink]\nreader.dmLevel=melspec\nfilename=%s\n" % (os.path.join(d, "m.htk"), os.path.join(d, "m.csv")))
            c2 = os.path.join(d, "m.conf")
            open(c2, "w").write(txt)
            r = subprocess.run([refrun.SMILEXTRACT, "-C", c2, "-I", wav, "-O", os.path.join(d, "x.htk"), "-l", "1"] + opts, capture_output=True, text=True)
            if r.returncode:
                print(r.stderr[-2000:])
                sys.exit(1)
            rows, _ = refrun.read_htk(os.path.join(d, "m.htk"))
            out["melspec_" + name] = rows.astype(np.float32)
            out["melspec_names"] = np.array(open(os.path.join(d, "m.csv")).readline().strip().split(";")[2:])
            print("melspec", name, rows.shape, float(np.abs(rows).max()), out["melspec_names"][:2])
    out["variants"] = np.array(list(VARIANTS))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "melspec_scales.npz"), **out)


if __name__ == "__main__":
    main()
