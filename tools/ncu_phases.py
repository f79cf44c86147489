#!/usr/bin/env python3
"""Instruction / stall-sample share per marked source region of an ncu report.
usage: python tools/ncu_phases.py report.ncu-rep file.cuh  (regions = lines containing '// ----' or '// ====' markers)"""
import csv, io, subprocess, sys
rep, src = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur, agg = None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) >= 8 and r[0].isdigit():
        try:
            agg.append((cur, int(r[0]), int(r[7]), int(r[4])))
        except ValueError:
            pass
tot = sum(a[2] for a in agg) or 1
smp = sum(a[3] for a in agg) or 1
base = src.split("/")[-1]
marks = [(1, "(file head)")]
for n, line in enumerate(open(src), 1):
    t = line.strip()
    if t.startswith("// ----") or t.startswith("// ====") or t.startswith("// B:") or t.startswith("// C:"):
        marks.append((n, t[:70]))
marks.append((10**9, ""))
for (a, name), (b, _) in zip(marks, marks[1:]):
    i = sum(x[2] for x in agg if x[0] == base and a <= x[1] < b)
    s = sum(x[3] for x in agg if x[0] == base and a <= x[1] < b)
    if i:
        print(f"{a:4d} {i / 1e9:6.2f}G {100 * i / tot:5.1f}%  samples {100 * s / smp:5.1f}%  {name}")
oth = {}
for x in agg:
    if x[0] != base:
        o = oth.setdefault(x[0], [0, 0]); o[0] += x[2]; o[1] += x[3]
for k, v in oth.items():
    print(f"     {v[0] / 1e9:6.2f}G {100 * v[0] / tot:5.1f}%  samples {100 * v[1] / smp:5.1f}%  {k}")
print(f"total {tot / 1e9:.2f} G warp-instructions, {smp} samples")
