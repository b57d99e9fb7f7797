#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` export by SASS opcode: executed warp instructions and
stall samples per opcode, plus the hottest instructions.  Usage: sass_mix.py file.csv [configs]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
n_cfg = float(sys.argv[2]) if len(sys.argv) > 2 else None
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
ex, st = defaultdict(float), defaultdict(float)
tot_ex = tot_st = 0.0
items = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    src = r[ci["Source"]].strip()
    op = src.split()[0] if src else "?"
    if op.startswith("@"):
        op = src.split()[1]
    op = op.split(".")[0]
    e = float(r[ci["Instructions Executed"]] or 0)
    s = float(r[ci["Warp Stall Sampling (All Samples)"]] or 0)
    ex[op] += e
    st[op] += s
    tot_ex += e
    tot_st += s
    items.append((s, e, src))
print("total warp instructions executed: %.4g   stall samples: %.0f" % (tot_ex, tot_st))
if n_cfg:
    print("thread instructions per configuration: %.0f" % (tot_ex * 32 / n_cfg))
print("%-10s %12s %7s %9s %7s" % ("opcode", "executed", "%", "samples", "%"))
for op in sorted(ex, key=lambda k: -ex[k])[:28]:
    print("%-10s %12.4g %6.1f%% %9.0f %6.1f%%" % (op, ex[op], 100 * ex[op] / tot_ex, st[op], 100 * st[op] / max(tot_st, 1)))
print("\nhottest instructions by stall samples:")
for s, e, src in sorted(items, reverse=True)[:25]:
    print("%7.0f %10.4g  %s" % (s, e, src[:110]))
