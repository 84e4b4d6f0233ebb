#!/bin/bash
mkdir -p gpurun_out
for sp in 1 0; do
VELA_B200_JIT_SPECULATE=$sp timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 30 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('spec=$sp bench', l['ms_per_step'], l['roofline']['stage_ms'])"
VELA_B200_JIT_SPECULATE=$sp timeout 300 python tools/probes/gpu_stage_probe.py config3_gp 10000 8192 2>&1 | tail -1 | cut -c1-140 | sed "s/^/spec=$sp probe /"
done > gpurun_out/b2.log 2>&1
cat gpurun_out/b2.log
