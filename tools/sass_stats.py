#!/usr/bin/env python
"""SASS statistics of the per-model (NVRTC) chain kernel, without a GPU.

    python tools/sass_stats.py config3_gp [--ntoa 2000] [--lines] [--kernel vela_chain_spec]

Builds the synthetic model with the CPU oracle as the phase evaluator (tests/helpers.py), asks the library to
compile the specialised translation unit (vela_b200_debug_spec_compile; VELA_B200_JIT_DUMP keeps the cubin),
disassembles it with nvdisasm and counts instructions per opcode class and, with --lines, per source line.
"""
import argparse
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

FP64 = ("DADD", "DMUL", "DFMA", "DSETP", "DMNMX", "MUFU", "DMMA", "F2F", "I2F", "F2I")


def classify(op):
    base = op.split(".")[0]
    if base in ("DADD", "DMUL", "DFMA", "DSETP", "DMMA"):
        return base
    if base in ("MUFU",):
        return "MUFU"
    if base in ("SHFL",):
        return "SHFL"
    if base in ("LDS", "STS", "LDSM"):
        return base
    if base in ("LDG", "STG", "LD", "ST", "LDC", "LDL", "STL", "ULDC", "LDCU"):
        return base
    if base in ("BRA", "BSSY", "BSYNC", "CALL", "RET", "EXIT", "WARPSYNC", "BAR", "BRX", "JMP"):
        return "CTRL"
    return "OTHER"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config")
    ap.add_argument("--ntoa", type=int, default=None)
    ap.add_argument("--lines", action="store_true")
    ap.add_argument("--kernel", default="vela_chain_spec")
    ap.add_argument("--keep", default=None, help="keep the cubin at this path")
    ap.add_argument("--top", type=int, default=40)
    args = ap.parse_args()

    from __graft_entry__ import load_package
    vb = load_package()
    import oracle as O
    from helpers import oracle_eval_factory
    cfg = vb.synth.CONFIGS[args.config]
    ntoa = args.ntoa or min(cfg["ntoa"], 2000)
    ds = vb.synth.make_dataset(args.config, oracle_eval_factory(vb, O), ntoa=ntoa)
    md = vb.ModelDescriptor(ds.model, ds.toas)
    cubin = args.keep or os.path.join(tempfile.gettempdir(), f"vela_spec_{args.config}.cubin")
    os.environ["VELA_B200_JIT_DUMP"] = cubin
    L = vb.lib.lib()
    rc = L.vela_b200_debug_spec_compile(md.byref())
    msg = L.vela_b200_last_error().decode()
    if rc <= 0:
        raise SystemExit(f"compile failed: {msg}")
    for ln in msg.splitlines():
        if "registers" in ln or "spill" in ln or "Compiling entry" in ln:
            print(ln.strip())
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    # split per function
    cur = None
    per_op = collections.Counter()
    per_line = collections.defaultdict(collections.Counter)
    line_tag = "?"
    in_kernel = False
    for ln in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            in_kernel = m.group(1) == args.kernel
            continue
        if not in_kernel:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            inl = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
            line_tag = f"{os.path.basename(m.group(1))}:{m.group(2)}"
            if inl:
                line_tag += " <- " + " <- ".join(f"{os.path.basename(f)}:{l}" for f, l in inl[:3])
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
        if m:
            op = m.group(2)
            c = classify(op)
            per_op[c] += 1
            per_line[line_tag][c] += 1
    total = sum(per_op.values())
    print(f"{args.kernel}: {total} SASS instructions")
    for k, v in per_op.most_common():
        print(f"  {k:8s} {v}")
    fp64 = sum(per_op[k] for k in ("DADD", "DMUL", "DFMA", "DSETP", "MUFU"))
    print(f"  FP64-pipe (DADD+DMUL+DFMA+DSETP+MUFU): {fp64}")
    if args.lines:
        rows = sorted(per_line.items(), key=lambda kv: -sum(kv[1].values()))
        for tag, cnt in rows[: args.top]:
            n = sum(cnt.values())
            d = sum(cnt[k] for k in ("DADD", "DMUL", "DFMA", "DSETP", "MUFU"))
            print(f"{n:5d} (fp64 {d:4d}, shfl {cnt['SHFL']:3d})  {tag}")


if __name__ == "__main__":
    main()
