#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_synth.py tests/test_gpu_fixtures.py tests/test_gpu_jit.py -m gpu -q --maxfail=15 > gpurun_out/c2_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/c2_tests.log
tail -3 gpurun_out/c2_tests.log
timeout 300 python tools/probes/gpu_chain_seed_probe.py config3_gp 10000 8192 > gpurun_out/c2_seed.log 2>&1; cat gpurun_out/c2_seed.log
timeout 300 python tools/probes/gpu_stage_probe.py config2_ell1 5000 4096 2>&1 | tail -1
