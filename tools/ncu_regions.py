#!/usr/bin/env python
"""Warp-instructions / stall samples of an ncu source page per named source region.
usage: python tools/ncu_regions.py x.csv envs  file:lo-hi=name ...   (x.csv from ncu --page source --csv --print-source cuda,sass)"""
import collections, csv, sys
path, envs = sys.argv[1], float(sys.argv[2])
regions = []
for a in sys.argv[3:]:
    spec, name = a.split("=")
    f, rng = spec.split(":")
    lo, hi = rng.split("-")
    regions.append((f, int(lo), int(hi), name))
rows = list(csv.reader(open(path)))
cur_file, hdr = None, None
agg = collections.OrderedDict()
for r in rows:
    if len(r) >= 2 and r[0] in ("File Path", "File Name"):
        cur_file = r[1].split("/")[-1]
    elif len(r) > 8 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) > 8 and r[0] != "" and r[2] == "-":
        d = dict(zip(hdr, r))
        try:
            n = int(d["Instructions Executed"])
        except ValueError:
            continue
        ln = int(r[0])
        name = "other:" + cur_file
        for f, lo, hi, nm in regions:
            if f == cur_file and lo <= ln <= hi:
                name = nm
                break
        a = agg.setdefault(name, [0, 0, 0, 0])
        a[0] += n; a[1] += int(d["# Samples"] or 0); a[2] += int(d.get("L1 Wavefronts Shared") or 0); a[3] += int(d.get("L1 Wavefronts Shared Excessive") or 0)
tot = sum(a[0] for a in agg.values()); samp = sum(a[1] for a in agg.values())
print("total %.0f warp-instr/env, %d samples" % (tot / envs, samp))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-28s %7.1f instr/env %5.1f%%   samples %5.1f%%   smem wavefronts/env %7.1f (excess %6.1f)" % (k, a[0] / envs, 100.0 * a[0] / tot, 100.0 * a[1] / max(samp, 1), a[2] / envs, a[3] / envs))
