#!/usr/bin/env python
"""Opcode histogram per kernel phase (or for one source line) of an ncu capture.
usage: ncu_ops.py <report> <mangled-kernel-substring> [--line N] [--frames NFRAMES] [--src file.cu]"""
import collections, csv, re, subprocess, sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines as nl
rep, ksub = sys.argv[1], sys.argv[2]
line = int(sys.argv[sys.argv.index("--line") + 1]) if "--line" in sys.argv else None
frames = float(sys.argv[sys.argv.index("--frames") + 1]) if "--frames" in sys.argv else 1e6
so = os.path.join(nl.ROOT, "opensmile_b200", "libosm_b200.so")
for a in sys.argv[3:]:
    if a.endswith(".so"):
        so = a
table = nl.line_table(so, ksub)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
hdr, body = rows[hi], rows[hi + 1:]
iI = hdr.index("Instructions Executed")
fns = {k[0] for k in table}
fn = next((f for f in fns if sum(1 for k in table if k[0] == f) == len(body)), None)   # the instance that was profiled
if fn is None:
    fn = list(fns)[0]
    print("warning: no instance with %d instructions, using %s" % (len(body), fn), file=sys.stderr)
offs = sorted(k[1] for k in table if k[0] == fn)
srcname = sys.argv[sys.argv.index("--src") + 1] if "--src" in sys.argv else "kernels.cu"
marks = []
for i, ln in enumerate(open(os.path.join(nl.ROOT, "opensmile_b200", "csrc", srcname)).read().splitlines(), 1):
    m = re.search(r"// =================\s*(.*?)\s*=*$", ln)
    if m:
        marks.append((i, m.group(1)[:28]))
ph = collections.defaultdict(collections.Counter)
for idx, r in enumerate(body):
    if idx >= len(offs):
        break
    outer, inn, sass = table[(fn, offs[idx])]
    if line is not None:
        if outer[1] != line:
            continue
        name = "line %d" % line
    else:
        name = "(other)"
        if outer[0].endswith(srcname):
            for i, nm in marks:
                if outer[1] >= i:
                    name = nm
    n = int(r[iI]); t = sass.split(); op = t[1] if t[0].startswith("@") else t[0]
    ph[name][op] += n
for name, ops in ph.items():
    tot = sum(ops.values())
    print("%-30s %.1f/frame: " % (name, tot / frames) + ", ".join("%s %.1f" % (op, n / frames) for op, n in ops.most_common(18)))
