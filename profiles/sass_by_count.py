#!/usr/bin/env python
"""Group the SASS of one kernel (ncu --page source --csv) by execution count per configuration: the straight-line phase 1
runs once, the box loop S/4 x B times, the per-sphere part S times ...  Prints, per class, the static instruction
count, the executed thread instructions per configuration and the opcode mix.  Usage: sass_by_count.py file.csv configs"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
n_cfg = float(sys.argv[2])
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
ci = {h: i for i, h in enumerate(rows[hi])}
classes = defaultdict(lambda: [0, 0.0, 0.0, defaultdict(float)])
tot = 0.0
for r in rows[hi + 1:]:
    if len(r) < len(rows[hi]):
        continue
    src = r[ci["Source"]].strip()
    op = src.split()[0] if src else "?"
    if op.startswith("@"):
        op = src.split()[1]
    full = op
    op = op.split(".")[0]
    if full.startswith("IMAD.MOV") or full.startswith("IMAD.IADD") or full.startswith("IMAD.WIDE") or full.startswith("IMAD.SHL"):
        op = ".".join(full.split(".")[:2])
    e = float(r[ci["Instructions Executed"]] or 0)
    s = float(r[ci["Warp Stall Sampling (All Samples)"]] or 0)
    per = e * 32 / n_cfg
    key = round(per, 1) if per < 3 else round(per)
    c = classes[key]
    c[0] += 1; c[1] += per; c[2] += s; c[3][op] += per
    tot += per
print("thread instructions per configuration: %.0f" % tot)
for key in sorted(classes, key=lambda k: -classes[k][1])[:14]:
    n, per, s, ops = classes[key]
    mix = ", ".join("%s %.0f" % (o, v) for o, v in sorted(ops.items(), key=lambda kv: -kv[1])[:12])
    print("x%-6s static %5d  executed/cfg %7.0f (%4.1f%%)  samples %6.0f | %s" % (key, n, per, 100 * per / tot, s, mix))
