#!/bin/bash
# Round 2, GPU call P: rank-300 panel Cholesky with the row-wise assembly and 16-byte panel IO.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_edge_cases.py -m gpu -q --maxfail=15 > gpurun_out/p_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/p_tests.log
tail -4 gpurun_out/p_tests.log
for cfg in "config5_array 20000 2048" "config5_array 100000 8192"; do
  timeout 300 python tools/probes/gpu_stage_probe.py $cfg 2>&1 | tail -1
done > gpurun_out/p_stages.log 2>&1
cat gpurun_out/p_stages.log
