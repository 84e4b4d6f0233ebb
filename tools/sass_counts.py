#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of the shipped library (and, optionally, of per-model NVRTC cubins).

    python tools/sass_counts.py [extra.cubin ...] > profiles/r2_sass_counts.txt

Evidence for which hardware paths the kernels use: DMMA (FP64 tensor cores, mma.sync.m8n8k4.f64), LDGSTS (cp.async),
SYNCS (mbarrier), DFMA/DADD/DMUL (FP64 pipe), SHFL, UTCMMA/LDTM/UTMALDG (tcgen05 / TMEM / TMA - absent: no FP64 kind).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["DMMA", "DFMA", "DADD", "DMUL", "MUFU", "LDGSTS", "SYNCS", "SHFL", "LDS", "STS", "LDG", "STG", "BAR",
        "UTCMMA", "LDTM", "UTMALDG", "UBLKCP"]


def counts(path):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for ln in txt.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
        if m and cur:
            per[cur][m.group(2)] += 1
            per[cur]["_total"] += 1
    return per


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
    except Exception:
        return n


def main():
    paths = [os.path.join(ROOT, "vela.jl_b200", "libvela_b200.so")] + sys.argv[1:]
    for p in paths:
        print(f"== {os.path.relpath(p, ROOT) if p.startswith(ROOT) else p}")
        arch = subprocess.run(["cuobjdump", "-lelf", p], capture_output=True, text=True).stdout.strip()
        print("   " + arch.replace("\n", "; "))
        print("   " + "kernel".ljust(70) + " total " + " ".join(k.rjust(6) for k in KEYS))
        for fn, c in counts(p).items():
            name = re.sub(r"\(.*", "", demangle(fn))[:68]
            print("   " + name.ljust(70) + f"{c['_total']:6d} " + " ".join(str(c[k]).rjust(6) for k in KEYS))


if __name__ == "__main__":
    main()
