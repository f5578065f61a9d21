#!/usr/bin/env python
"""Turns the ncu artefacts a gpurun call brought back (gpurun_out/) into the committed
summaries under profiles/:
  python scripts/summarize_profiles.py <tag> <launches.csv> <full.ncu-rep> [<early-state.ncu-rep>]"""
import collections
import csv
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag, launches, rep = sys.argv[1], Path(sys.argv[2]), Path(sys.argv[3])
out = ROOT / "profiles"
out.mkdir(exist_ok=True)

# -- launch list (gpu__time_duration.sum, cold-cache and serialised: compare SHARES) ------
rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
hdr, data = rows[0], rows[1:]
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
with open(out / f"{tag}_launches.csv", "w") as f:
    f.write("id,kernel,grid,block,duration_ns\n")
    for r in data:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("unnamed>::", "")
        ns = float(r[ix["Metric Value"]].replace(",", ""))
        f.write(f'{r[ix["ID"]]},{name},{r[ix["Grid Size"]].replace(",", " ")},{r[ix["Block Size"]].replace(",", " ")},{ns:.0f}\n')
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns / 1e6
tot = sum(a[1] for a in agg.values())
lines = [f"# {tag}: ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu` (16M grid, one B200)",
         "# ncu --metrics gpu__time_duration.sum --clock-control none; per-launch times are cold-cache and",
         "# serialised, so only the SHARE of the step is comparable with bench.py's CUDA-event stage times.",
         "", "| kernel | launches | total ms | share | avg ms |", "|---|---|---|---|---|"]
for k, (c, t) in agg.items():
    lines.append(f"| {k} | {c} | {t:.3f} | {100 * t / tot:.1f} % | {t / c:.4f} |")

# -- full captures: one plainly launched step ----------------------------------------------
def full_table(rep, heading):
    global h2, units, d2, jx, names, lines
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h2, units, d2 = rr[0], rr[1], rr[2:]
    jx = {h: i for i, h in enumerate(h2)}
    want = [("gpu__time_duration.sum", "duration"), ("launch__registers_per_thread", "regs/thread"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
            ("sm__inst_executed.avg.per_cycle_active", "IPC (max ~3.6 measured, 4 nominal)"),
            ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
            ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
            ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
            ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
            ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / instr"),
            ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
            ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
            ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
            ("smsp__inst_executed.sum", "warp instructions")]
    names = [r[jx["Kernel Name"]].split("(")[0].replace("void ", "").replace("unnamed>::", "") for r in d2]
    lines += ["", f"# {tag}: ncu --set full --clock-control none, ONE plainly launched step of the 16M grid workload at the",
              heading,
              "", "| metric | " + " | ".join(names) + " |", "|---|" + "---|" * len(names)]
    for key, label in want:
        if key in jx:
            vals = []
            for r in d2:
                v = r[jx[key]]
                try:
                    v = f"{float(v.replace(',', '')):.4g}"
                except ValueError:
                    pass
                vals.append(f"{v} {units[jx[key]]}".strip())
            lines.append(f"| {label} | " + " | ".join(vals) + " |")
    stalls = [h for h in h2 if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
    lines += ["", "Warp stall reasons (warps stalled per issue; > 0.5 only):", "", "| reason | " + " | ".join(names) + " |",
              "|---|" + "---|" * len(names)]
    for hname in stalls:
        vals = [float(r[jx[hname]] or 0) for r in d2]
        if max(vals) > 0.5:
            short = hname.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")
            lines.append(f"| {short} | " + " | ".join(f"{v:.2f}" for v in vals) + " |")


early = Path(sys.argv[4]) if len(sys.argv) > 4 else None

import json


def _bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def write_traffic(suffix, what):
    """per-launch DRAM traffic of every kernel of the capture just tabulated (bench.py reports it as
    roofline.traffic for the run whose state matches)"""
    traffic = {}
    for r, nm in zip(d2, names):
        rd = _bytes(r[jx["dram__bytes_read.sum"]], units[jx["dram__bytes_read.sum"]])
        wr = _bytes(r[jx["dram__bytes_write.sum"]], units[jx["dram__bytes_write.sum"]])
        traffic.setdefault(nm, []).append(rd + wr)
    (out / f"{tag}_traffic_{suffix}.json").write_text(json.dumps(
        {"source": f"profiles/{tag}_ncu_summary.md: ncu --set full --clock-control none, one plainly launched step of the "
                   f"16M grid workload {what}; dram__bytes_read.sum + dram__bytes_write.sum per launch",
         "bytes_per_launch": {k: sum(v) / len(v) for k, v in traffic.items()}}, indent=1) + "\n")


full_table(rep, "# state after 100 steps (scripts/profile_step.py --pre 100, the floor pile-up): mean candidates 116, mean neighbours 25.")
write_traffic("late", "after 100 steps (floor pile-up, mean candidates 116)")
if early is not None:
    full_table(early, "# state after 3 steps (--pre 3, the undisturbed lattice: the sparse regime): mean candidates 39, mean neighbours 7.")
    write_traffic("early", "after 3 steps (undisturbed lattice, mean candidates 39)")
(out / f"{tag}_ncu_summary.md").write_text("\n".join(lines) + "\n")
print("\n".join(lines))
