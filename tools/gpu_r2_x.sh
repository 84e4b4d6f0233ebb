#!/bin/bash
# Round 2, GPU call X: panel Cholesky with the DMMA row solve (inverse of the diagonal block) and look-ahead.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_synth.py -m gpu -q --maxfail=15 > gpurun_out/x_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/x_tests.log
tail -4 gpurun_out/x_tests.log
for cfg in "config5_array 20000 2048" "config5_array 100000 8192"; do
  timeout 300 python tools/probes/gpu_stage_probe.py $cfg 2>&1 | tail -1
done > gpurun_out/x_stages.log 2>&1
cat gpurun_out/x_stages.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:chol_panel -c 1 -o gpurun_out/x_panel -f python tools/probes/gpu_stage_probe.py config5_array 20000 2048 > gpurun_out/x_ncu.log 2>&1; echo "ncu rc $?"
