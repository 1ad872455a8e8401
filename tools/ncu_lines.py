#!/usr/bin/env python
"""Summarise `ncu --page source --csv --print-source cuda,sass` output per CUDA source line.

usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass --kernel-id :::N > x.csv
       python tools/ncu_lines.py x.csv [top] [envs]
"""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
envs = float(sys.argv[3]) if len(sys.argv) > 3 else 131072.0
rows = list(csv.reader(open(path)))
cur_file, hdr, agg = None, None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif len(r) > 8 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[2] == "-":   # per-line aggregate (Address '-')
        d = dict(zip(hdr, r))
        try:
            n = int(d["Instructions Executed"])
        except ValueError:
            continue
        agg.append((n, int(d["# Samples"] or 0), cur_file, r[0], r[1].strip()[:100],
                    int(d.get("L1 Wavefronts Shared Excessive") or 0)))
tot = sum(a[0] for a in agg)
samp = sum(a[1] for a in agg)
print("total warp-instructions %d  (%.0f per env)  samples %d" % (tot, tot / envs, samp))
for n, s, f, ln, src, exc in sorted(agg, reverse=True)[:top]:
    print("%6.2f%% inst %6.2f%% samp  excess-wf %9d  %s:%s  %s" % (100.0 * n / tot, 100.0 * s / max(samp, 1), exc, f, ln, src))
