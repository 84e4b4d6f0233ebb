#!/usr/bin/env python
"""Dynamic instruction mix and stall hot spots of one kernel from an `ncu --set full --import-source on` report.

    python tools/ncu_source.py report.ncu-rep kernel_regex [--top 30] [--units N]

Reads `ncu -i report --page source --csv`: per SASS instruction the executed warp-instruction count and the stall samples.
--units N divides the counts by N (e.g. warps x walkers) to give instructions per unit of work.
"""
import argparse
import collections
import csv
import io
import re
import subprocess


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("kernel")
    ap.add_argument("--top", type=int, default=30)
    ap.add_argument("--units", type=float, default=1.0)
    ap.add_argument("--dump", action="store_true", help="print every executed SASS line with its count and samples")
    a = ap.parse_args()
    txt = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv", "--kernel-name", f"regex:{a.kernel}"],
                         capture_output=True, text=True).stdout
    lines = txt.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
    per = collections.Counter()
    samp = collections.Counter()
    tot = 0
    tot_s = 0
    out = []
    for r in rows:
        try:
            n = int(r["Instructions Executed"])
            s = int(r["Warp Stall Sampling (All Samples)"])
        except (ValueError, KeyError, TypeError):
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", r["Source"])
        op = m.group(2) if m else "?"
        per[op] += n
        samp[op] += s
        tot += n
        tot_s += s
        out.append((n, s, r["Source"].strip()))
    print(f"total warp instructions {tot}  ({tot / a.units:.1f} per unit), stall samples {tot_s}")
    for op, n in per.most_common(25):
        print(f"  {op:10s} {n / a.units:9.2f} per unit  {100.0 * n / tot:5.1f}% of instructions  {100.0 * samp[op] / max(tot_s, 1):5.1f}% of samples")
    print("-- top stall lines")
    for n, s, src in sorted(out, key=lambda t: -t[1])[: a.top]:
        print(f"  {100.0 * s / max(tot_s, 1):5.2f}%  exec {n / a.units:8.2f}  {src[:110]}")
    if a.dump:
        print("-- all executed lines")
        for n, s, src in out:
            if n:
                print(f"  {n / a.units:8.2f} {s:6d}  {src[:120]}")


if __name__ == "__main__":
    main()
