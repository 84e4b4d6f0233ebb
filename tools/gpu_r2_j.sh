#!/bin/bash
# Round 2, GPU call J: fast-arithmetic build of the chain (bit-mirrored unit vector without sqrt/divisions, fused small terms).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=15 > gpurun_out/j_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/j_tests.log
tail -4 gpurun_out/j_tests.log
for cfg in "config3_gp 10000 8192" "config4_dd_wb 20000 8192" "config2_ell1 5000 4096"; do
  timeout 300 python tools/probes/gpu_stage_probe.py $cfg 2>&1 | tail -1
  VELA_B200_STRICT=1 timeout 300 python tools/probes/gpu_stage_probe.py $cfg 2>&1 | tail -1 | sed 's/^/  [strict] /'
done > gpurun_out/j_stages.log 2>&1
cat gpurun_out/j_stages.log
timeout 600 python tools/probes/gpu_variants_probe.py config3_gp 10000 8192 default,mb4,unroll2_mb5,unroll2_mb4,unroll2_mb3,t256_mb2,t64_mb10 > gpurun_out/j_var_c3.log 2>&1
cat gpurun_out/j_var_c3.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err; echo "bench rc $?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/j_bench.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step')}, l['e2e']['value'], l.get('parity'), l['roofline']['stage_ms'])
PY
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -o gpurun_out/j_c3 -f python bench.py --profile --no-extras --no-cpu-baseline > gpurun_out/j_ncu_c3.log 2>&1; echo "ncu c3 rc $?"
