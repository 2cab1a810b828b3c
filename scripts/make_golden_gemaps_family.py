"""Generate tests/golden/gemaps_family.npz + gemaps_headers.json with the UNMODIFIED reference (oracle/_ref/SMILExtract):

    python scripts/make_golden_gemaps_family.py        # needs `make -C oracle ref` (build container only)

The five shipped feature-set files of the GeMAPS family with -lldcsvoutput on mixed_pcm(24000, seed=3): element names and
row counts (json) and the rows themselves (npz, key = file name without extension)."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refrun  # noqa: E402
from opensmile_b200.synth import mixed_pcm  # noqa: E402

CONFS = ("gemaps/v01a/GeMAPSv01a.conf", "gemaps/v01b/GeMAPSv01b.conf", "egemaps/v01a/eGeMAPSv01a.conf",
         "egemaps/v01b/eGeMAPSv01b.conf", "egemaps/v02/eGeMAPSv02.conf")


def main():
    assert refrun.available(), "build the reference first: make -C oracle ref"
    pcm = mixed_pcm(24000, 16000, seed=3)
    hdr, rows = {}, {}
    for conf in CONFS:
        with tempfile.TemporaryDirectory() as d:
            refrun.write_wav(os.path.join(d, "in.wav"), pcm, 16000, 1)
            subprocess.run([refrun.SMILEXTRACT, "-C", os.path.join(refrun.CONFIG_DIR, conf), "-I", os.path.join(d, "in.wav"),
                            "-lldcsvoutput", os.path.join(d, "o.csv"), "-lldhtkoutput", os.path.join(d, "o.htk"), "-l", "0"],
                           check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            lines = open(os.path.join(d, "o.csv")).read().splitlines()
            r = refrun.read_htk(os.path.join(d, "o.htk"))[0]
        hdr[conf] = {"names": lines[0].split(";")[2:], "rows_m24k": len(lines) - 1}
        rows[os.path.splitext(os.path.basename(conf))[0]] = r
    json.dump(hdr, open(os.path.join(ROOT, "tests", "golden", "gemaps_headers.json"), "w"), indent=1)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "gemaps_family.npz"), **rows)
    print({k: v.shape for k, v in rows.items()})


if __name__ == "__main__":
    main()
