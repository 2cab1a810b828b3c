import os, subprocess, sys, tempfile
import numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/scripts")
from oracle import refrun
from opensmile_b200.synth import mixed_pcm, voiced_pcm
LEVELS = ["gemapsv01b_lld_single_logF0_smo", "gemapsv01b_loudness_smo", "egemapsv02_lldSetNoF0AndLoudnessZ_smo", "egemapsv02_lldSetNoF0AndLoudnessNz_smo",
          "egemapsv02_lldSetSpectralNz_smo", "egemapsv02_lldSetSpectralZ_smo", "egemapsv02_energyRMS"]
rec = np.load("/root/repo/tests/golden/egemaps_recordings.npz")
sigs = {"m24k": mixed_pcm(24000, 16000, seed=3), "v32k": voiced_pcm(32000, 16000, seed=7), "rec": rec["pcm_opensmile_16k"]}
out = {}
for key, pcm in sigs.items():
    with tempfile.TemporaryDirectory() as d:
        wav = os.path.join(d, "in.wav")
        refrun.write_wav(wav, pcm, 16000, 1)
        conf = os.path.join(d, "w.conf")
        txt = open("/root/reference/config/egemaps/v02/eGeMAPSv02.conf").read()
        txt = txt.replace("\\{../../", "\\{/root/reference/config/").replace("\\{eGeMAPSv02_core", "\\{/root/reference/config/egemaps/v02/eGeMAPSv02_core")
        txt += "\n[componentInstances:cComponentManager]\n" + "".join("instance[dbg%d].type=cCsvSink\n" % i for i in range(len(LEVELS)))
        for i, l in enumerate(LEVELS):
            txt += "[dbg%d:cCsvSink]\nreader.dmLevel=%s\nfilename=%s\nappend=0\ntimestamp=0\nnumber=0\nprintHeader=1\n" % (i, l, os.path.join(d, "l%d.csv" % i))
        open(conf, "w").write(txt)
        r = subprocess.run([refrun.SMILEXTRACT, "-C", conf, "-I", wav, "-csvoutput", os.path.join(d, "f.csv"), "-l", "1"], capture_output=True, text=True)
        if r.returncode: print(r.stderr[-2000:]); sys.exit(1)
        for i, l in enumerate(LEVELS):
            lines = open(os.path.join(d, "l%d.csv" % i)).read().strip().split("\n")
            hdr = lines[0].split(";")
            rows = np.array([[float(x) for x in ln.split(";")] for ln in lines[1:]], np.float32) if len(lines) > 1 else np.zeros((0, len(hdr)), np.float32)
            out["%s_%s" % (key, l)] = rows
            out["names_" + l] = np.array(hdr)
            print(key, l, rows.shape)
np.savez_compressed("/root/repo/tests/golden/gemaps_func_levels.npz", **out)
