#!/bin/bash
# Round 2, GPU call I: single-launch stage 2 (unit table + split-K): parity tests, stage times by batch size.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_synth.py tests/test_gpu_full_size.py tests/test_gpu_fixtures.py -m gpu -q --maxfail=15 > gpurun_out/i_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/i_tests.log
tail -4 gpurun_out/i_tests.log
for cfg in "config3_gp 10000 8192" "config3_gp 10000 4096" "config3_gp 10000 2048" "config3_gp 10000 1024" "config4_dd_wb 20000 8192" "config5_array 100000 8192"; do
  VELA_B200_TWO_STAGE_MIN=1 timeout 300 python tools/probes/gpu_stage_probe.py $cfg 2>&1 | tail -1
  VELA_B200_TWO_STAGE=0 timeout 300 python tools/probes/gpu_stage_probe.py $cfg 2>&1 | tail -1 | sed 's/^/  [two-stage off] /'
done > gpurun_out/i_stages.log 2>&1
cat gpurun_out/i_stages.log
for ks in 1 2 3 4; do VELA_B200_STAGE2_KSPLIT=$ks timeout 300 python tools/probes/gpu_stage_probe.py config3_gp 10000 8192 2>&1 | tail -1 | sed "s/^/  [stage2 ks $ks] /"; done | tee gpurun_out/i_ks.log
