#!/usr/bin/env python3
"""Executed-instruction histogram by opcode for the kernels of an ncu report matching a regex (--import-source on):
   python tools/ncu_opcode_hist.py report.ncu-rep <kernel regex> [top N]"""
import csv
import subprocess
import sys
from collections import Counter

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
c, cs = Counter(), Counter()
hdr = None
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    elif hdr and len(r) > iE and r[iE].isdigit():
        t = r[iS].split()
        op = t[1] if t[0].startswith("@") else t[0]
        c[op] += int(r[iE])
        cs[op] += int(r[iSm]) if r[iSm].isdigit() else 0
tot, ts = sum(c.values()), max(sum(cs.values()), 1)
print("%s: %d executed warp instructions" % (rx, tot))
for op, n in c.most_common(top):
    print("  %-24s %6.2f%% inst %6.2f%% smp" % (op, 100 * n / tot, 100 * cs[op] / ts))
