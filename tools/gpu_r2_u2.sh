#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_jit.py -m gpu -q --maxfail=15 > gpurun_out/u2_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/u2_tests.log
tail -15 gpurun_out/u2_tests.log
