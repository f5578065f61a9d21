#!/usr/bin/env python
"""Per-instruction execution counts / stall samples of one kernel from `ncu --page source --csv`,
printed as a compact listing: offset, executed warp-instr per warp (÷ warps), samples, SASS."""
import csv, sys
path, warps = sys.argv[1], float(sys.argv[2])
rows = list(csv.reader(open(path)))
hdr = rows[1]
ix = {n: i for i, n in enumerate(hdr)}
tot_i = tot_s = 0
out = []
base = None
for r in rows[2:]:
    if r and r[0] == "Kernel Name": break      # next kernel instance
    if len(r) < len(hdr) - 5 or r[0] == "Address": continue
    addr = int(r[ix["Address"]], 16)
    if base is None: base = addr
    ie = float(r[ix["Instructions Executed"]] or 0)
    sm = float(r[ix["# Samples"]] or 0)
    th = float(r[ix["Avg. Threads Executed"]] or 0)
    l1 = r[ix["L1 Tag Requests Global"]]
    tot_i += ie; tot_s += sm
    out.append((addr - base, ie / warps, sm, th, l1, r[ix["Source"]].strip()))
print(f"total executed per warp {tot_i / warps:.1f}, samples {tot_s:.0f}")
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
for o, ie, sm, th, l1, src in out:
    if lo <= o <= hi:
        print(f"{o:05x} {ie:7.2f} {100 * sm / tot_s:5.2f}% {th:5.1f} {l1:>9s}  {src}")
