#!/bin/bash
# Round 2, GPU call D: tests with the graph path / slab reduce / new sincos_cr, latency probe, variants, sanitizer on a small case.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=15 > gpurun_out/d_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/d_tests.log
tail -4 gpurun_out/d_tests.log
timeout 300 python tools/probes/gpu_latency_probe.py > gpurun_out/d_lat.log 2>&1; tail -8 gpurun_out/d_lat.log
VELA_B200_GRAPHS=0 timeout 300 python tools/probes/gpu_latency_probe.py > gpurun_out/d_lat_nograph.log 2>&1; tail -8 gpurun_out/d_lat_nograph.log | head -4
timeout 600 python tools/probes/gpu_variants_probe.py config3_gp 10000 8192 default,chol_cta > gpurun_out/d_var_c3.log 2>&1; cat gpurun_out/d_var_c3.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_fixtures.py -m gpu -q -x > gpurun_out/d_memcheck.log 2>&1; tail -5 gpurun_out/d_memcheck.log
