"""Per-column parity table of a shipped feature-set configuration on the GPU against the reference's LLD rows.

    python scripts/parity_report.py [out.md]        (GPU box; reads tests/golden/*.npz only)

For every golden signal the rows of the CUDA path (Session.extract_pcm through the C ABI) are compared column by column with
the rows the unmodified reference wrote for the same PCM (-lldhtkoutput, exact float32).  Error = |got - ref| relative to
the column's largest magnitude in the reference rows.  Columns: max, median, share of rows beyond 1e-5 and beyond 1e-3.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opensmile_b200.session import Session  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REFCONF = os.path.join(ROOT, "oracle", "_ref", "config")


def cases():
    """(title, conf, options, [(label, pcm int16, sample_rate, n_chan, ref rows)])"""
    from opensmile_b200.synth import mixed_pcm
    G = np.load(os.path.join(GOLD, "formant_goldens.npz"))
    sig = [("mixed_pcm(24000, seed 3)", mixed_pcm(24000, 16000, seed=3), 16000, 1, "m24k"),
           ("mixed_pcm(40000, seed 5)", mixed_pcm(40000, 16000, seed=5), 16000, 1, "m40k")]
    out = [("eGeMAPSv02.conf (BASELINE configs[2])", "egemaps/v02/eGeMAPSv02.conf", {"lldcsvoutput": "x.csv"},
            [(t, x, sr, nc, G["egemaps_lld_" + k]) for t, x, sr, nc, k in sig]),
           ("GeMAPSv01b.conf", "gemaps/v01b/GeMAPSv01b.conf", {"lldcsvoutput": "x.csv"},
            [(t, x, sr, nc, G["gemaps_lld_" + k]) for t, x, sr, nc, k in sig])]
    p = os.path.join(GOLD, "egemaps_recordings.npz")            # scripts/make_golden_recordings.py
    if os.path.exists(p):
        R = np.load(p)
        rec = []
        for key in sorted(k[4:] for k in R.files if k.startswith("pcm_")):
            rec.append((key + " (%d Hz)" % int(R["sr_" + key]), R["pcm_" + key], int(R["sr_" + key]), 1, R["egemaps_" + key]))
        out[0][3].extend(rec)
        if any(k.startswith("compare_") for k in R.files):
            out.append(("ComParE_2016.conf (BASELINE configs[3])", "compare16/ComParE_2016.conf", {"lldcsvoutput": "x.csv"},
                        [(key + " (%d Hz)" % int(R["sr_" + key]), R["pcm_" + key], int(R["sr_" + key]), 1, R["compare_" + key])
                         for key in sorted(k[4:] for k in R.files if k.startswith("pcm_")) if "compare_" + key in R.files]))
    return out


def column_table(names, got, ref):
    scale = np.abs(ref).max(axis=0) + 1e-30
    err = np.abs(got - ref) / scale
    rows = []
    for j, n in enumerate(names):
        e = err[:, j]
        rows.append((n, float(e.max()), float(np.median(e)), float((e > 1e-5).mean() * 100), float((e > 1e-3).mean() * 100)))
    return rows


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "parity_report.md")
    lines = ["# Per-column parity of the CUDA path against the reference's LLD rows", "",
             "error = |got - ref| / max_t |ref[t, column]|; reference rows = tests/golden (unmodified reference, -lldhtkoutput)", ""]
    summary = {}
    for title, conf, opts, sigs in cases():
        for label, pcm, sr, nc, ref in sigs:
            lines += ["## %s -- %s: %d rows x %d columns" % (title, label, ref.shape[0], ref.shape[1]), ""]
            try:
                s = Session(os.path.join(REFCONF, conf), options=opts, device=0)
                names = s.element_names(float(sr), nc)
                rows, fo = s.extract_pcm(np.concatenate([pcm, np.zeros(8 * nc, np.int16)]), np.array([0, len(pcm) // nc], np.int64), float(sr), nc)
                s.close()
            except Exception as e:                                   # an unsupported format is reported, not hidden
                lines += ["NOT RUN: %s" % e, ""]
                summary[title + " / " + label] = "not run: %s" % e
                continue
            if rows.shape != ref.shape:
                lines += ["SHAPE MISMATCH got %s ref %s" % (rows.shape, ref.shape), ""]
                summary[title + " / " + label] = "shape mismatch"
                continue
            tab = column_table(names, rows, ref)
            lines += ["| column | max | median | % rows > 1e-5 | % rows > 1e-3 |", "|---|---|---|---|---|"]
            for n, mx, md, p5, p3 in tab:
                lines.append("| %s | %.2e | %.2e | %.1f | %.1f |" % (n, mx, md, p5, p3))
            lines.append("")
            summary[title + " / " + label] = {"worst_max": max(t[1] for t in tab), "columns_le_1e-5": sum(t[1] <= 1e-5 for t in tab),
                                              "columns": len(tab)}
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    open(out_path, "w").write("\n".join(lines) + "\n")
    print(json.dumps(summary, indent=1))


if __name__ == "__main__":
    main()
