#!/usr/bin/env python
"""Key counters of every kernel in an .ncu-rep, one column per kernel (read here, not on the GPU box)."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, units, data = rr[0], rr[1], rr[2:]
jx = {n: i for i, n in enumerate(h)}
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio" ]
stall = [n for n in h if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio")]
for r in data:
    print("==", r[jx["Kernel Name"]][:70], "grid", r[jx["Grid Size"]])
    for w in want:
        if w in jx:
            print(f"   {w:75s} {r[jx[w]]:>14s} {units[jx[w]]}")
    ss = sorted(((float(r[jx[n]] or 0), n) for n in stall), reverse=True)[:7]
    for v, n in ss:
        print(f"   stall {n[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:30s} {v:.2f}")
