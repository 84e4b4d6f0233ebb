#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1200 python -m pytest tests -m gpu -q --maxfail=15 > gpurun_out/s2_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/s2_tests.log
tail -3 gpurun_out/s2_tests.log
timeout 600 python bench.py --steps 30 --no-extras 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(l['value'], l['ms_per_step'], l['e2e']['value'], l['roofline']['frac'], l['roofline']['kernels'][0].get('pipe_active_ncu'), l['clocks'])"
