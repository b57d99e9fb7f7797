#!/usr/bin/env python
"""Per-CUDA-source-line totals from `ncu --page source --csv --print-source cuda,sass`:
warp instructions executed and stall samples per line.  Usage: line_mix.py file.csv [configs]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n_cfg = float(sys.argv[2]) if len(sys.argv) > 2 else None
out = []
for r in rows:
    if len(r) > 8 and r[0].isdigit():
        try:
            out.append((int(r[0]), float(r[7] or 0), float(r[4] or 0), r[1].strip()))
        except ValueError:
            pass
tot_e = sum(o[1] for o in out)
tot_s = sum(o[2] for o in out)
print("total executed %.4g, samples %.0f" % (tot_e, tot_s))
for line, e, s, src in sorted(out, key=lambda o: -o[1])[:45]:
    per = " %7.0f/cfg" % (e * 32 / n_cfg) if n_cfg else ""
    print("%4d %6.2f%% ex %6.2f%% st%s  %s" % (line, 100 * e / tot_e, 100 * s / tot_s, per, src[:100]))
