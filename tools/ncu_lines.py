#!/usr/bin/env python
"""Attribute `ncu --page source --csv` stall samples (per SASS instruction) to source lines.

  python tools/ncu_lines.py <source.csv> <cubin> <kernel-name-substring> [N]

The CSV carries absolute instruction addresses; `nvdisasm -g` of the cubin gives, for the same function, the
(file, line) of every instruction offset.  Samples are summed per source line and the N hottest lines printed with their
dominant stall reasons.  The cubin comes from `cuobjdump -xelf all climt_b200/libclimt_b200.so` (compile with -lineinfo)."""
import collections
import csv
import re
import subprocess
import sys

path, cubin, kname = sys.argv[1], sys.argv[2], sys.argv[3]
N = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(path, newline="")))
hi = next(i for i, r in enumerate(rows) if any("Sampl" in c for c in r))
h = rows[hi]
body = [r for r in rows[hi + 1:] if len(r) == len(h)]


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


ai = h.index("Address")
si = next(i for i, c in enumerate(h) if "Sampling" in c and "All" in c)
stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
base = min(int(r[ai], 16) for r in body)
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
line_of, cur, infn = {}, None, False
for ln in dis:
    if ln.startswith(".text.") or re.match(r"^\s*\.section\s+\.text\.", ln):
        infn = kname in ln
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
    m = re.match(r"^\s*/\*([0-9a-f]{4,})\*/", ln)
    if m and infn:
        line_of[int(m.group(1), 16)] = cur
per = collections.defaultdict(lambda: [0.0, collections.Counter(), 0.0])
tot = 0.0
for r in body:
    off = int(r[ai], 16) - base
    key = line_of.get(off, ("?", 0))
    s = num(r[si])
    per[key][0] += s
    per[key][2] += num(r[h.index("Instructions Executed")])
    for i, c in stall_cols:
        per[key][1][c] += num(r[i])
    tot += s
print(f"{tot:.0f} samples, {len(body)} instructions, {len(line_of)} offsets with line info")
for key, (s, st, ins) in sorted(per.items(), key=lambda kv: -kv[1][0])[:N]:
    top = ", ".join(f"{k[6:]} {100 * v / max(s, 1):.0f}%" for k, v in st.most_common(3))
    print(f"{100 * s / tot:5.1f}%  {key[0]}:{key[1]:<5d} inst {ins:10.0f}   {top}")
