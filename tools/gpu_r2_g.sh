#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_synth.py tests/test_gpu_full_size.py -m gpu -q --maxfail=10 > gpurun_out/g_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/g_tests.log
tail -4 gpurun_out/g_tests.log
timeout 300 python tools/probes/gpu_stage_probe.py config5_array 100000 8192 > gpurun_out/g_c5_new.log 2>&1; tail -1 gpurun_out/g_c5_new.log
timeout 300 python tools/probes/gpu_stage_probe.py config5_array 20000 2048 > gpurun_out/g_c5_small.log 2>&1; tail -1 gpurun_out/g_c5_small.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:chol_panel -c 1 -o gpurun_out/g_panel -f python tools/probes/gpu_stage_probe.py config5_array 20000 2048 > gpurun_out/g_ncu.log 2>&1; echo "ncu rc $?"
