#!/bin/bash
# per-kernel average durations (ncu launch list) of a general plan: kernel_times.sh <conf> [n_utt] [n_samples]
ncu --metrics gpu__time_duration.sum --clock-control none --csv python scripts/bench_general.py "$1" "${2:-300}" "${3:-48000}" 2>/dev/null | python -c '
import csv,sys,collections
rows=list(csv.reader(sys.stdin))
hi=next(i for i,r in enumerate(rows) if "Kernel Name" in r)
h=rows[hi]; kn=h.index("Kernel Name"); mv=h.index("Metric Value")
agg=collections.OrderedDict(); cnt=collections.Counter()
for r in rows[hi+1:]:
    if len(r)<=mv: continue
    n=r[kn]
    if "osm::" not in n: continue
    n=n.split("(")[0][:60]
    agg[n]=agg.get(n,0)+float(r[mv].replace(",","")); cnt[n]+=1
for n,v in agg.items(): print("%-62s n=%3d avg %.1f us" % (n,cnt[n],v/cnt[n]/1000))
'
