#!/bin/bash
# Round 2, GPU call H: two-stage column sums - parity tests, stage times on/off for configs 3/4/5.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=15 > gpurun_out/h_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/h_tests.log
tail -4 gpurun_out/h_tests.log
for cfg in "config3_gp 10000 8192" "config4_dd_wb 20000 8192" "config5_array 100000 8192" "config3_gp 10000 1024"; do
  timeout 300 python tools/probes/gpu_stage_probe.py $cfg 2>&1 | tail -1
  VELA_B200_TWO_STAGE=0 timeout 300 python tools/probes/gpu_stage_probe.py $cfg 2>&1 | tail -1 | sed 's/^/  [two-stage off] /'
done > gpurun_out/h_stages.log 2>&1
cat gpurun_out/h_stages.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err; echo "bench rc $?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/h_bench.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step')}, l['e2e']['value'], l.get('parity'), l['roofline']['stage_ms'])
PY
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -o gpurun_out/h_c3 -f python bench.py --profile --no-extras --no-cpu-baseline > gpurun_out/h_ncu_c3.log 2>&1; echo "ncu c3 rc $?"
