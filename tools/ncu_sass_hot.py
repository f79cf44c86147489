#!/usr/bin/env python3
"""Hot SASS regions of an ncu report: executed warp-instructions and stall samples per block of consecutive
instructions (split at branch targets is not attempted: fixed windows of N instructions).
usage: python tools/ncu_sass_hot.py report.ncu-rep [window=40] [top=25]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
win = int(sys.argv[2]) if len(sys.argv) > 2 else 40
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ia, isrc, ismp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
ins = [(r[isrc].strip(), int(r[iex]), int(r[ismp])) for r in rows[2:] if len(r) > iex and r[iex].isdigit()]
tot = sum(i[1] for i in ins); smp = sum(i[2] for i in ins)
print(f"{len(ins)} SASS instructions, {tot/1e9:.2f} G executed, {smp} samples")
blocks = []
for b in range(0, len(ins), win):
    chunk = ins[b:b + win]
    blocks.append((b, sum(c[1] for c in chunk), sum(c[2] for c in chunk), chunk))
for b, ex, sm, chunk in sorted(blocks, key=lambda x: -x[1])[:top]:
    ops = {}
    for c in chunk:
        op = c[0].split()[0] if not c[0].startswith("@") else c[0].split()[1]
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + c[1]
    topops = ", ".join(f"{k} {v/ex*100:.0f}%" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:6])
    print(f"[{b:5d}..{b+win:5d}) {ex/1e9:6.2f}G {100*ex/tot:5.1f}%  samples {100*sm/smp:5.1f}%  first: {chunk[0][0][:40]:40s} | {topops}")
