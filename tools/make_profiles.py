#!/usr/bin/env python
"""Turn an `ncu --set full` report of the headline step into the committed summaries under profiles/:

    python tools/make_profiles.py gpurun_out/<report>.ncu-rep profiles/r2_ncu_full_config3_<tag>.txt [profiles/r2_ncu_config3.json]

The .txt is the per-kernel counter summary (profiles/ncu_summarise.py keys); the .json holds what bench.py reads for its
`roofline` block: per kernel the DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) and the busy
fraction of the bounding pipe (sm__pipe_fp64_cycles_active / sm__pipe_tensor_cycles_active, pct_of_peak_sustained_active).
"""
import csv
import io
import json
import subprocess
import sys

rep, txt_out = sys.argv[1], sys.argv[2]
json_out = sys.argv[3] if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
sys.path.insert(0, "profiles")
keys = None
for line in open("profiles/ncu_summarise.py"):
    if line.startswith("keys="):
        keys = eval(line.split("=", 1)[1])
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
out = []
kernels = {}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d["Kernel Name"]
    out.append("---- " + name[:100])
    for k in keys:
        if k in d and d[k] not in ("", "0"):
            out.append(f"   {k} {d[k]} {units[hdr.index(k)]}")
    short = name.split("(")[0].replace("void ", "").replace("vela::", "").split("<")[0]
    rd = float(d["dram__bytes_read.sum"]) * scale.get(units[hdr.index("dram__bytes_read.sum")], 1.0)
    wr = float(d["dram__bytes_write.sum"]) * scale.get(units[hdr.index("dram__bytes_write.sum")], 1.0)
    fp64 = float(d.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", 0) or 0)
    tens = float(d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0) or 0)
    dur = float(d["gpu__time_duration.sum"]) * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(units[hdr.index("gpu__time_duration.sum")], 1.0)
    ent = {"dram_bytes": rd + wr, "pipe_active": max(fp64, tens) / 100.0, "pipe": "tensor (DMMA)" if tens > fp64 else "fp64",
           "ncu_duration_ms": dur, "warps_active_pct": float(d.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0) or 0),
           "issue_active_pct": float(d.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0) or 0),
           "inst_executed": float(d.get("smsp__inst_executed.sum", 0) or 0), "launches": 1}
    if short in kernels:   # several launches of one kernel in the step (the two stages of the column sums): bytes, time and
        old = kernels[short]   # instructions add up, the percentages are duration-weighted
        tot = old["ncu_duration_ms"] + dur
        for k in ("pipe_active", "warps_active_pct", "issue_active_pct"):
            ent[k] = (old[k] * old["ncu_duration_ms"] + ent[k] * dur) / tot
        for k in ("dram_bytes", "ncu_duration_ms", "inst_executed", "launches"):
            ent[k] += old[k]
    kernels[short] = ent
open(txt_out, "w").write("\n".join(out) + "\n")
if json_out:
    json.dump({"source": txt_out, "command": "ncu --profile-from-start off --set full --import-source on --clock-control none "
               "python bench.py --profile --no-extras --no-cpu-baseline", "kernels": kernels}, open(json_out, "w"), indent=1)
print("\n".join(out))
