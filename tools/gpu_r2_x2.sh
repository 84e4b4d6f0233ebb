#!/bin/bash
# Round 2, GPU call X2: solar-wind geometry by expansion (fast build): full GPU suite, config 4 and the others.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=15 > gpurun_out/x2_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/x2_tests.log
tail -3 gpurun_out/x2_tests.log
for cfg in "config2_ell1 5000 4096" "config3_gp 10000 8192" "config4_dd_wb 20000 8192" "config5_array 100000 8192"; do timeout 300 python tools/probes/gpu_stage_probe.py $cfg 2>&1 | tail -1 | cut -c1-125; done
VELA_B200_JIT_MINBLOCKS=6 timeout 300 python tools/probes/gpu_stage_probe.py config4_dd_wb 20000 8192 2>&1 | tail -1 | cut -c1-125 | sed 's/^/[mb6] /'
