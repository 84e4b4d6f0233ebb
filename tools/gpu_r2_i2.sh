#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_synth.py -m gpu -q --maxfail=15 > gpurun_out/i2_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/i2_tests.log
tail -3 gpurun_out/i2_tests.log
