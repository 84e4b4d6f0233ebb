#!/bin/bash
# quick perf probe: bench without the CPU baseline, prints stage times
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'stages', {k: round(v,3) for k,v in r['stage_ms'].items()})
for k in r['kernels']: print('  ', k['kernel'][:40], round(k['ms'],3), 'ms', round(k['achieved'],2), '/', round(k['peak'],2), 'TF frac', round(k['frac'],3))"
