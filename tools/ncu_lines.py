#!/usr/bin/env python
"""Summarise `ncu --page source --csv --print-source cuda,sass` output per CUDA source line.

usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > x.csv
       python tools/ncu_lines.py x.csv [top] [envs*launches] [inst|samp]
Lines that appear once per profiled launch are summed.
"""
import collections
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
envs = float(sys.argv[3]) if len(sys.argv) > 3 else 2 * 131072.0
key = sys.argv[4] if len(sys.argv) > 4 else "inst"
rows = list(csv.reader(open(path)))
cur_file, hdr = None, None
agg = collections.OrderedDict()
for r in rows:
    if len(r) >= 2 and r[0] in ("File Path", "File Name"):
        cur_file = r[1].split("/")[-1]
    elif len(r) > 8 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) > 8 and r[0] != "" and r[2] == "-":   # per-line aggregate (Address '-')
        d = dict(zip(hdr, r))
        try:
            n = int(d["Instructions Executed"])
        except ValueError:
            continue
        k = (cur_file, int(r[0]))
        a = agg.setdefault(k, [0, 0, 0, r[1].strip()[:110]])
        a[0] += n
        a[1] += int(d["# Samples"] or 0)
        a[2] += int(d.get("L1 Wavefronts Shared Excessive") or 0)
tot = sum(a[0] for a in agg.values())
samp = sum(a[1] for a in agg.values())
print("total warp-instructions %d  (%.0f per env)  samples %d" % (tot, tot / envs, samp))
idx = 0 if key == "inst" else 1
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][idx])[:top]:
    print("%6.2f%% inst %6.2f%% samp  excess-wf/env %6.1f  %s:%d  %s" % (
        100.0 * a[0] / tot, 100.0 * a[1] / max(samp, 1), a[2] / envs, f, ln, a[3]))
