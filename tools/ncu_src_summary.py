#!/usr/bin/env python
"""Summarise `ncu --page source --csv --print-source sass` output: instructions,
shared-memory wavefronts and stall samples per opcode, and the hottest SASS lines."""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {n: i for i, n in enumerate(hdr)}
tot_inst = tot_s = 0
by = defaultdict(lambda: [0, 0, 0, 0, 0])
lines = []
for r in rows[2:]:
    if len(r) < len(hdr) - 2 or r[ix["Instructions Executed"]] == "Instructions Executed":
        continue        # short rows and the repeated header of a second kernel
    op = r[ix["Source"]].split()
    if op and op[0].startswith("@"):
        op = op[1:]
    name = op[0].rstrip(";") if op else "?"
    inst = int(r[ix["Instructions Executed"]] or 0)
    samp = int(r[ix["# Samples"]] or 0)
    wf = int(r[ix["L1 Wavefronts Shared"]] or 0)
    ideal = int(r[ix["L1 Wavefronts Shared Ideal"]] or 0)
    b = by[name]
    b[0] += 1; b[1] += inst; b[2] += wf; b[3] += samp; b[4] += ideal
    tot_inst += inst; tot_s += samp
    lines.append((samp, inst, wf, ideal, r[ix["Source"]].strip(), r[ix["Address"]]))
print("total inst %d samples %d" % (tot_inst, tot_s))
for name, b in sorted(by.items(), key=lambda kv: -kv[1][1])[:24]:
    print("%-22s n=%4d inst=%12d %5.1f%% wf=%12d ideal=%12d samples=%8d %5.1f%%" % (
        name, b[0], b[1], 100.0 * b[1] / tot_inst, b[2], b[4], b[3], 100.0 * b[3] / max(tot_s, 1)))
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
agg = defaultdict(int)
for r in rows[2:]:
    if len(r) < len(hdr) - 2 or r[ix["Instructions Executed"]] == "Instructions Executed":
        continue
    for n in stalls:
        agg[n] += int(r[ix[n]] or 0)
print(sorted(((k, round(100.0 * v / max(tot_s, 1), 1)) for k, v in agg.items()), key=lambda kv: -kv[1])[:10])
if len(sys.argv) > 2:
    print("hottest lines:")
    for samp, inst, wf, ideal, src, addr in sorted(lines, reverse=True)[:int(sys.argv[2])]:
        print("%7d samp %11d inst wf %11d ideal %11d  %s" % (samp, inst, wf, ideal, src))
