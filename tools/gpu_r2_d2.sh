#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,launch__grid_size,launch__block_size,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/d2_small.csv python tools/probes/gpu_small_launches.py > gpurun_out/d2_ncu.log 2>&1; echo "rc $?"
python - <<'PY'
import csv
rows=list(csv.reader(l for l in open('gpurun_out/d2_small.csv') if l.startswith('"')))
hdr=rows[0]
for r in rows[1:]:
    d=dict(zip(hdr,r))
    if d['Metric Name']=='gpu__time_duration.sum': print(d['ID'], d['Kernel Name'][:70], d['Grid Size'], d['Block Size'], d['Metric Value'], d['Metric Unit'])
PY
