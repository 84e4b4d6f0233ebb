#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=15 > gpurun_out/h2_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/h2_tests.log
tail -3 gpurun_out/h2_tests.log
timeout 600 python tools/probes/gpu_latency_probe.py 2>&1 | head -4
for cfg in "config3_gp 10000 8192" "config4_dd_wb 20000 8192"; do timeout 300 python tools/probes/gpu_stage_probe.py $cfg 2>&1 | tail -1 | cut -c1-130; done
