#!/bin/bash
# Round 2, GPU call K: tensor-core register Cholesky (chol_dmma_kernel): parity, stage times on/off, latency, ncu.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_synth.py tests/test_gpu_edge_cases.py tests/test_gpu_fixtures.py -m gpu -q --maxfail=15 > gpurun_out/k_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/k_tests.log
tail -15 gpurun_out/k_tests.log
for cfg in "config3_gp 10000 8192" "config4_dd_wb 20000 8192" "config3_gp 10000 256" "config3_gp 10000 16"; do
  timeout 300 python tools/probes/gpu_stage_probe.py $cfg 2>&1 | tail -1
  VELA_B200_CHOL_DMMA=0 timeout 300 python tools/probes/gpu_stage_probe.py $cfg 2>&1 | tail -1 | sed 's/^/  [dmma off] /'
done > gpurun_out/k_stages.log 2>&1
cat gpurun_out/k_stages.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:chol_dmma -c 1 -o gpurun_out/k_dmma -f python tools/probes/gpu_stage_probe.py config3_gp 10000 8192 > gpurun_out/k_ncu.log 2>&1; echo "ncu rc $?"
