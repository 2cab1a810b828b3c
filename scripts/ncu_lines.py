#!/usr/bin/env python
"""Attribute an ncu capture's per-instruction counters to source lines / kernel phases.

usage: ncu_lines.py <report.ncu-rep> <kernel-substring> [lib.so] [--phases]

Joins `ncu --page source --csv` (SASS view, per-instruction counters) with the line table of the
cubin (`nvdisasm -gi`), keyed by instruction offset, and aggregates "Instructions Executed" and
stall samples by the OUTERMOST source line (the line inside the kernel body, inlined callees are
folded into their call site).  With --phases, lines are further folded into the phases marked in
kernels.cu by comments of the form  `// ================= name`.
"""
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_table(so, kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    table = {}
    for f in os.listdir(tmp):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur_fn, group, last_group = None, [], []
        for ln in txt.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                cur_fn = m.group(1)
                continue
            if cur_fn is None or kernel_sub not in cur_fn:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                group.append((m.group(1), int(m.group(2))))
                continue
            m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*);", ln)
            if m:
                if group:
                    last_group, group = group, []
                off = int(m.group(1), 16)
                table[(cur_fn, off)] = (last_group[-1] if last_group else ("?", 0), last_group[0] if last_group else ("?", 0), m.group(2).strip())
        if table:
            break
    return table


def main():
    rep, ksub = sys.argv[1], sys.argv[2]
    so = os.path.join(ROOT, "opensmile_b200", "libosm_b200.so")
    for a in sys.argv[3:]:
        if a.endswith(".so"):
            so = a
    phases = "--phases" in sys.argv
    table = line_table(so, ksub)
    fns = {k[0] for k in table}
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hi = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
    hdr = rows[hi]
    iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = [(h, i) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    # choose the function whose instruction count matches
    body = rows[hi + 1:]
    fn = None
    for f in fns:
        if sum(1 for k in table if k[0] == f) == len(body):
            fn = f
    if fn is None:
        fn = max(fns, key=lambda f: -abs(sum(1 for k in table if k[0] == f) - len(body)))
        print("warning: instruction count mismatch (%d in report); using %s" % (len(body), fn), file=sys.stderr)
    offs = sorted(k[1] for k in table if k[0] == fn)
    agg = {}
    tot_i = tot_s = 0
    for idx, r in enumerate(body):
        if idx >= len(offs):
            break
        outer, inner, sass = table[(fn, offs[idx])]
        try:
            n, s = int(r[iI]), int(r[iS])
        except ValueError:
            continue
        st = {}
        for h, i in stall_cols:
            try:
                st[h] = int(r[i])
            except ValueError:
                pass
        key = outer
        a = agg.setdefault(key, {"inst": 0, "samp": 0, "stalls": {}})
        a["inst"] += n; a["samp"] += s
        for h, v in st.items():
            a["stalls"][h] = a["stalls"].get(h, 0) + v
        tot_i += n; tot_s += s
    src_cache = {}

    def src(f, l):
        if f not in src_cache:
            try:
                src_cache[f] = open(f).read().splitlines()
            except OSError:
                src_cache[f] = []
        L = src_cache[f]
        return L[l - 1].strip() if 0 < l <= len(L) else ""

    if phases:
        # the kernel's own source file: the one most outer lines belong to (kernels.cu, lld_fast.cu, ...)
        cnt = {}
        for (f, l), a in agg.items():
            if f.endswith(".cu"):
                cnt[f] = cnt.get(f, 0) + a["inst"]
        kfile = max(cnt, key=cnt.get) if cnt else None
        marks = []
        if kfile:
            for i, ln in enumerate(open(kfile).read().splitlines(), 1):
                m = re.search(r"// =================\s*(.*?)\s*=*$", ln)
                if m:
                    marks.append((i, m.group(1)))
        ph = {}
        for (f, l), a in agg.items():
            name = "(other)"
            if f == kfile:
                for i, nm in marks:
                    if l >= i:
                        name = nm
            p = ph.setdefault(name, {"inst": 0, "samp": 0, "stalls": {}})
            p["inst"] += a["inst"]; p["samp"] += a["samp"]
            for h, v in a["stalls"].items():
                p["stalls"][h] = p["stalls"].get(h, 0) + v
        print("%-60s %8s %8s  top stalls" % ("phase", "inst%", "samp%"))
        for nm, p in sorted(ph.items(), key=lambda kv: -kv[1]["samp"]):
            top = sorted(p["stalls"].items(), key=lambda kv: -kv[1])[:4]
            print("%-60s %7.2f%% %7.2f%%  %s" % (nm[:60], 100.0 * p["inst"] / tot_i, 100.0 * p["samp"] / max(tot_s, 1),
                                                 ", ".join("%s=%d" % (h[6:], v) for h, v in top)))
    else:
        print("total warp-instructions %d, samples %d" % (tot_i, tot_s))
        for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1]["samp"])[:45]:
            top = sorted(a["stalls"].items(), key=lambda kv: -kv[1])[:3]
            print("%6.2f%% inst %6.2f%% samp  %s:%d  %-70s %s" % (100.0 * a["inst"] / tot_i, 100.0 * a["samp"] / max(tot_s, 1),
                  os.path.basename(f), l, src(f, l)[:70], ", ".join("%s=%d" % (h[6:], v) for h, v in top)))


if __name__ == "__main__":
    main()
