#!/usr/bin/env python
"""SASS of the hot kernels of the in-tree libsph_b200.so -> profiles/r02_sass_hot_kernels.txt
(cuobjdump -sass, instruction lines only).  CPU only.

    python scripts/sass_listing.py
"""
import collections
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "cudafluidsimulator_b200" / "libsph_b200.so"
WANT = [  # (substring of the mangled name, title)
    ("k_density_flatILb0ELb1ELb0E", "k_density_flat<0,1,0> (production: factored sum, same predicate)"),
    ("k_force_integrate_flatILb1E", "k_force_integrate_flat<1> (with the fused per-cell count: ATOMG at the end)"),
    ("k_reorderILb1E", "k_reorder<1> (counting sort: rank inside the cell by index, then the gather)"),
    ("k_cell_scan_apply", "k_cell_scan_apply (cell_start = exclusive prefix of the counts)"),
    ("k_cell_scatter", "k_cell_scatter"),
    ("k_cell_count", "k_cell_count (stand-alone count; warp-aggregated atomics)"),
    ("k_onesweepILb0E", "k_onesweep<0> (radix arm)"),
    ("k_density_tileILb1ELb0E", "k_density_tile<1,0> (TMA-staged dense tiles, opt-in)"),
    ("k_msg_flags", "k_msg_flags (peer-memory hand-shake)"),
    ("k_ghost_install", "k_ghost_install"),
]
sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
funcs, name = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = m.group(1)
        funcs[name] = []
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?;)\s*/\*", line)
    if m and name:
        funcs[name].append(f"{m.group(1)}  {m.group(2)}")
out = ["# r02: SASS of the hot kernels of libsph_b200.so (cuobjdump -sass, sm_100a), instruction lines only;",
       "# regenerate with scripts/sass_listing.py.",
       "# What to look for: FADD2 / FMUL2 / FFMA2 (Blackwell packed f32x2 math) and SHF.L.W (one funnel shift per",
       "# candidate outcome) in the density pair loop; ATOMG + SHFL at the end of the force kernel (fused count);",
       "# UBLKCP + SYNCS (1-D bulk TMA + mbarrier) in k_density_tile; LD/ST with .SYS scope and no MEMBAR in k_msg_flags;",
       "# LDG.E.128 gathers in the force and reorder kernels.", ""]
for key, title in WANT:
    hit = [n for n in funcs if key in n]
    if not hit:
        out += [f"## {title}: not found", ""]
        continue
    n = hit[0]
    ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", l.split("  ", 1)[1]).split(".")[0].split()[0] for l in funcs[n])
    out += [f"## {title}", f"## {n[:150]}",
            f"## {len(funcs[n])} instructions; most frequent opcodes: " + ", ".join(f"{k} {v}" for k, v in ops.most_common(16))]
    out += funcs[n] + [""]
(ROOT / "profiles" / "r02_sass_hot_kernels.txt").write_text("\n".join(out) + "\n")
print("\n".join(l for l in out if l.startswith("## ")))
