#!/bin/bash
# quick perf probe: bench without the CPU baseline, prints stage times
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'stages', {k: round(v,3) for k,v in r['stage_ms'].items()}, 'sigma TF', round(r['achieved'],2), 'peak', round(r['peak'],2), 'frac', round(r['frac'],3))"
