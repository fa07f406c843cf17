#!/usr/bin/env python3
"""Summarise an .ncu-rep (ncu --set full) into one CSV row per kernel launch: the metrics DESIGN.md / profiles/ quote.
  python tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.csv]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
        "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_xu.sum", "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_lsu.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct",
        "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
        "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio", "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio",
        "smsp__average_warp_latency_issue_stalled_wait.ratio", "smsp__average_warp_latency_issue_stalled_not_selected.ratio",
        "smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio", "smsp__average_warp_latency_issue_stalled_no_instruction.ratio"]


def main():
    rep = sys.argv[1]
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-units", "base"], text=True)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = ["ID", "Kernel Name", "Grid Size", "Block Size"] + [k for k in KEYS if k in idx]
    out = io.StringIO()
    w = csv.writer(out)
    w.writerow(cols)
    w.writerow([units[idx[c]] for c in cols])
    for r in rows[2:]:
        w.writerow([r[idx[c]] for c in cols])
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(out.getvalue())
    for r in rows[2:]:
        print("----", r[idx["Kernel Name"]].split("(")[0], r[idx["Grid Size"]])
        for k in KEYS:
            if k in idx:
                print("   %-72s %14s %s" % (k, r[idx[k]], units[idx[k]]))


if __name__ == "__main__":
    main()
