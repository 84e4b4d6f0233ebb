#!/bin/bash
mkdir -p gpurun_out
for cfg in "config2_ell1 5000 4096" "config3_gp 10000 8192" "config4_dd_wb 20000 8192" "config5_array 100000 8192"; do timeout 300 python tools/probes/gpu_stage_probe.py $cfg 2>&1 | tail -1 | cut -c1-125; done
timeout 1200 python -m pytest tests -m gpu -q --maxfail=15 > gpurun_out/y2_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/y2_tests.log
tail -3 gpurun_out/y2_tests.log
timeout 300 python tools/probes/gpu_latency_probe2.py 16 64 256 1024 2>&1
