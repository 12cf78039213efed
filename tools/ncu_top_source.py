#!/usr/bin/env python
"""Reduce `ncu --page source --csv` output to the N hottest source lines (by warp-stall samples)."""
import csv
import sys

path, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 60
rows = list(csv.reader(open(path, newline="")))
hdr_i = next(i for i, r in enumerate(rows) if any("Sampl" in c for c in r))
hdr = rows[hdr_i]
col = next(i for i, c in enumerate(hdr) if "Sampling" in c and "All" in c) if any("Sampling" in c and "All" in c for c in hdr) else \
    next(i for i, c in enumerate(hdr) if "Sampl" in c)
body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


body.sort(key=lambda r: -num(r[col]))
tot = sum(num(r[col]) for r in body) or 1.0
w = csv.writer(sys.stdout)
w.writerow(["share_%"] + hdr)
for r in body[:n]:
    w.writerow([f"{100 * num(r[col]) / tot:.2f}"] + r)
