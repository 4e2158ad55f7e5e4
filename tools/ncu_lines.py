#!/usr/bin/env python
"""Per-CUDA-source-line executed-instruction and stall-sample counts of one kernel from an `ncu --set full --import-source on`
report.  python tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]; ii = hdr.index("Instructions Executed"); si = hdr.index("# Samples")
L = []
for r in rows[hi + 1:]:
    if len(r) <= ii or r[0] == "":
        continue
    try:
        L.append((int(r[0]), int(r[ii]), int(r[si]), r[1].strip()[:100]))
    except ValueError:
        pass
tot = sum(x[1] for x in L); samp = sum(x[2] for x in L)
grid = None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
try:
    grid = float(rr[2][rr[0].index("launch__grid_size")])
except Exception:
    grid = 1.0
print(f"{rows[1][1][:90] if len(rows) > 1 else ''}\ntotal warp-instructions {tot}  ({tot / grid:.0f} per CTA, grid {grid:.0f}), stall samples {samp}")
for ln, n, s, src in sorted(L, key=lambda x: -x[1])[:top]:
    print(f"{n / grid:8.1f}/CTA  {100.0 * s / max(1, samp):5.1f}% smp  L{ln:<5d} {src}")
