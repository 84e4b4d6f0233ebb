#!/usr/bin/env python
"""Stall samples and executed instructions of one kernel of an ncu report, cut at its block barriers:

    python tools/ncu_segments.py report.ncu-rep kernel_substring [instances-for-normalisation]

One line per barrier-delimited segment of the SASS (share of the warp-stall samples, warp instructions per instance) with
the source line range it maps to, so that the phases of a multi-phase kernel can be weighed against each other."""
import csv, io, subprocess, sys
rep, sub = sys.argv[1], sys.argv[2]
norm = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
st = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and sub in r[1]][0]
hdr = rows[st + 1]
body = []
for r in rows[st + 2:]:
    if r and r[0] == "Kernel Name":
        break
    body.append(r)
iS, iE, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = sum(int(r[iN]) for r in body if len(r) > iN)
toti = sum(int(r[iE]) for r in body if len(r) > iN)
print(f"{rows[st][1]}: {len(body)} SASS instructions, {tot} samples, {toti / norm / 1e3:.1f} k warp instructions per instance")
start = cur = curi = 0
ops = {}
for idx, r in enumerate(body):
    if len(r) <= iN:
        continue
    cur += int(r[iN]); curi += int(r[iE])
    op = r[iS].split()
    op = (op[1] if op and op[0].startswith("@") else (op[0] if op else "?")).split(".")[0]
    ops[op] = ops.get(op, 0) + int(r[iE])
    if "BAR.SYNC" in r[iS] or "EXIT" in r[iS]:
        if cur > 0.002 * tot:
            top = ", ".join(f"{k} {v / norm / 1e3:.1f}k" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:5])
            print(f"  sass {start:5d}-{idx:5d}: samples {100 * cur / tot:5.1f} %  inst {curi / norm / 1e3:7.1f} k   [{top}]")
        start, cur, curi, ops = idx + 1, 0, 0, {}
