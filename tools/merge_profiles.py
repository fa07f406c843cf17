#!/usr/bin/env python3
"""Merges the ncu --set full captures of tools/profile_r1d.sh (gpurun_out/prof_r1d*.ncu-rep) into
profiles/<tag>_ncu_full_summary.csv (one row per kernel of one pass, base units) and profiles/traffic.json (DRAM bytes per
pass and stage = the `roofline.traffic` of the bench line).   python tools/merge_profiles.py r1d"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1d"
reps = [os.path.join(ROOT, "gpurun_out", "prof_%s%s.ncu-rep" % (tag, s)) for s in ("", "_2", "_3", "_f")]
rows_all, hdr, units = [], None, None
for rep in reps:
    if not os.path.exists(rep):
        continue
    tmp = "/tmp/%s.csv" % os.path.basename(rep)
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, tmp], stdout=subprocess.DEVNULL)
    rows = list(csv.reader(open(tmp)))
    if hdr is None:
        hdr, units = rows[0], rows[1]
    assert rows[0] == hdr and rows[1] == units, rep
    rows_all += [r for r in rows[2:] if "at::" not in r[1]]
ix = {h: i for i, h in enumerate(hdr)}
out, seen = [], set()
for r in rows_all:
    name = r[ix["Kernel Name"]].split("(")[0].replace("sdorb::", "").replace("void ", "")
    key = (name, r[ix["Grid Size"]])
    if key in seen:  # the same launch captured twice (second pass / second report)
        continue
    seen.add(key)
    out.append(r)
w = csv.writer(open(os.path.join(ROOT, "profiles", tag + "_ncu_full_summary.csv"), "w"))
w.writerow(hdr)
w.writerow(units)
w.writerows(out)
traffic = {}
for r in out:
    name = r[ix["Kernel Name"]].split("(")[0].replace("sdorb::", "").replace("void ", "").split("<")[0]
    name = {"resize_level_pre_kernel": "resize_level_kernel"}.get(name, name)
    t = traffic.setdefault(name, {"dram_bytes_per_launch": 0.0, "kernels_per_pass": 0, "us_per_pass_under_ncu": 0.0})
    t["dram_bytes_per_launch"] += float(r[ix["dram__bytes_read.sum"]]) + float(r[ix["dram__bytes_write.sum"]])
    t["kernels_per_pass"] += 1
    t["us_per_pass_under_ncu"] += float(r[ix["gpu__time_duration.sum"]]) / 1e3
    if "smsp__inst_executed.sum" in ix:
        t["warp_inst_per_launch"] = t.get("warp_inst_per_launch", 0.0) + float(r[ix["smsp__inst_executed.sum"]])
    # ncu --set full has no per-pipe instruction COUNT, only the pipe's utilisation: ALU-pipe warp-instructions of the launch =
    # utilisation x the pipe's issue rate (one warp-instruction per 2 clk per SM sub-partition = 2 / clk / SM: what the percentage
    # is of, and what tools/probe/pipe_probe2.cu and kernels_probe.cu measure) x the active SM cycles x 148 SMs.  Same for XU
    # (POPC: 16 lanes / clk / SM = 0.5 warp-instructions / clk / SM).
    if "sm__cycles_active.avg" in ix:
        cyc = float(r[ix["sm__cycles_active.avg"]])
        t["alu_warp_inst_per_launch"] = t.get("alu_warp_inst_per_launch", 0.0) + float(r[ix["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"]]) / 100 * 2.0 * cyc * 148
        t["xu_warp_inst_per_launch"] = t.get("xu_warp_inst_per_launch", 0.0) + float(r[ix["sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]]) / 100 * 0.5 * cyc * 148
    # busiest launch of the stage: what actually bounds it (none of these kernels waits on HBM)
    for key, col in (("alu_pipe_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                     ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                     ("dram_throughput_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")):
        t[key] = max(t.get(key, 0.0), float(r[ix[col]]))
for k, v in traffic.items():
    v["frames_per_pass"] = 256 if k.startswith("search") or k.startswith("match") else 512
    v["source"] = "profiles/%s_ncu_full_summary.csv (ncu --set full --clock-control none, one pass of 512 frames 640x480; tools/profile_%s.sh)" % (tag, tag[:2] if tag.startswith("r2") else tag)
    print("%-26s %d kernels %8.1f us %8.1f MB" % (k, v["kernels_per_pass"], v["us_per_pass_under_ncu"], v["dram_bytes_per_launch"] / 1e6))
json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
for r in out:
    g = lambda k: r[ix[k]]
    print("%-30s %-14s %8.1f us alu %5s fma %5s xu %5s issue %5s ipc %4s occ %5s regs %s" % (
        g("Kernel Name").split("(")[0].replace("sdorb::", "").replace("void ", "")[:30], g("Grid Size"), float(g("gpu__time_duration.sum")) / 1e3,
        g("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active")[:5], g("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active")[:5],
        g("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active")[:5], g("smsp__issue_active.avg.pct_of_peak_sustained_active")[:5],
        g("sm__inst_executed.avg.per_cycle_active")[:4], g("sm__warps_active.avg.pct_of_peak_sustained_active")[:5], g("launch__registers_per_thread")))
