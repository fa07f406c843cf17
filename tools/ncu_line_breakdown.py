#!/usr/bin/env python3
"""Executed warp instructions and stall samples per CUDA source line of one kernel of an ncu report (--import-source on):
   python tools/ncu_line_breakdown.py report.ncu-rep [top N]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, hdr, lines = "?", None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
        iE, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
    elif hdr and r[0].isdigit() and len(r) > iE and r[iE].isdigit():
        lines.append((fname, int(r[0]), r[1].strip(), int(r[iE]), int(r[iS]) if r[iS].isdigit() else 0))
tot, ts = sum(x[3] for x in lines), sum(x[4] for x in lines)
print("executed warp instructions %d, samples %d" % (tot, ts))
for f, n, src, e, sm in sorted(lines, key=lambda x: -x[3])[:top]:
    print("%5.2f%% inst %5.2f%% smp  %s:%d  %s" % (100 * e / tot, 100 * sm / ts, f, n, src[:110]))
