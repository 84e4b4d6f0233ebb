#!/bin/bash
# Round 2, GPU call M2: full GPU suite at HEAD (speculative chain, panel Cholesky), headline bench with extras, reference arm,
# ncu capture and launch list of the step.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=15 > gpurun_out/m2_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/m2_tests.log
tail -4 gpurun_out/m2_tests.log
timeout 600 python tools/probes/gpu_latency_probe.py > gpurun_out/m2_lat.log 2>&1; cat gpurun_out/m2_lat.log
SECONDS=0
timeout 1200 python bench.py > gpurun_out/m2_bench.json 2> gpurun_out/m2_bench.err; echo "bench rc $? in $SECONDS s"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/m2_bench.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step')}, l['e2e']['value'], l.get('parity'), l['roofline']['stage_ms'])
for k in ('other_configs','small_batch','dense_sigma_path'):
    print(k, json.dumps(l.get(k))[:1200])
PY
SECONDS=0
timeout 600 python bench.py --impl reference > gpurun_out/m2_ref.json 2> gpurun_out/m2_ref.err; echo "ref rc $? in $SECONDS s"; tail -c 300 gpurun_out/m2_ref.json
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -o gpurun_out/m2_c3 -f python bench.py --profile --no-extras --no-cpu-baseline > gpurun_out/m2_ncu_c3.log 2>&1; echo "ncu c3 rc $?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/m2_launches.csv python bench.py --profile --no-extras --no-cpu-baseline > gpurun_out/m2_ncu_l.log 2>&1; echo "ncu launches rc $?"
