#!/usr/bin/env python3
"""Per-phase and per-opcode breakdown of one kernel of an ncu report (captured with --import-source on):
   python tools/ncu_source_breakdown.py report.ncu-rep
Segments are the SASS ranges between BAR.SYNC instructions, in address order; shares are of executed warp instructions and of
warp-stall samples."""
import csv
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[start]
iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = [(r[iS].strip(), int(r[iE]), int(r[iSm])) for r in rows[start + 1:] if len(r) > iE and r[iE].isdigit()]
tot, ts = sum(d[1] for d in data), sum(d[2] for d in data)
print("kernel:", rows[0][1] if rows and len(rows[0]) > 1 else "?")
print("executed warp instructions %d, samples %d, SASS lines %d" % (tot, ts, len(data)))
seg, segs = 0, {}
for i, (s, e, sm) in enumerate(data):
    if "BAR.SYNC" in s:
        seg += 1
    v = segs.setdefault(seg, [0, 0, i, i])
    v[0] += e
    v[1] += sm
    v[3] = i
for k, v in segs.items():
    print("segment %d  inst %5.1f%%  samples %5.1f%%  SASS lines %d..%d" % (k, 100 * v[0] / tot, 100 * v[1] / ts, v[2], v[3]))
c, cs = Counter(), Counter()
for s, e, sm in data:
    t = s.split()
    op = t[1] if t[0].startswith("@") else t[0]
    c[op] += e
    cs[op] += sm
for op, n in c.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 24):
    print("%-22s %6.2f%% inst  %6.2f%% samples" % (op, 100 * n / tot, 100 * cs[op] / ts))
