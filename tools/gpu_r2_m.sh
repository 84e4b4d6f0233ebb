#!/bin/bash
# Round 2, GPU call M: fast chromatic powers (config 5 chain), stage-1 shape variants, synth tests.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_synth.py tests/test_gpu_full_size.py tests/test_gpu_jit.py -m gpu -q --maxfail=15 > gpurun_out/m_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/m_tests.log
tail -4 gpurun_out/m_tests.log
for sh in 0 1 2 3; do
  for cfg in "config3_gp 10000 8192" "config5_array 100000 8192"; do
    VELA_B200_STAGE1_SHAPE=$sh timeout 300 python tools/probes/gpu_stage_probe.py $cfg 2>&1 | tail -1 | sed "s/^/[stage1 shape $sh] /"
  done
done > gpurun_out/m_stages.log 2>&1
VELA_B200_STRICT=1 timeout 300 python tools/probes/gpu_stage_probe.py config5_array 100000 8192 2>&1 | tail -1 | sed "s/^/[strict] /" >> gpurun_out/m_stages.log
cat gpurun_out/m_stages.log
