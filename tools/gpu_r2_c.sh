#!/bin/bash
# Round 2, GPU call C: parity tests after the chain / Cholesky rewrites, stage probes, co-issue probe, ncu of the new kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=15 > gpurun_out/c_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/c_tests.log
tail -4 gpurun_out/c_tests.log
timeout 600 python tools/probes/gpu_variants_probe.py config3_gp 10000 8192 default,chol_cta,mb4,mb6,unroll2_mb5 > gpurun_out/c_var_c3.log 2>&1; cat gpurun_out/c_var_c3.log
timeout 300 python tools/probes/gpu_stage_probe.py config4_dd_wb 20000 8192 > gpurun_out/c_c4.log 2>&1; tail -1 gpurun_out/c_c4.log
timeout 300 python tools/probes/gpu_stage_probe.py config5_array 100000 8192 > gpurun_out/c_c5.log 2>&1; tail -1 gpurun_out/c_c5.log
timeout 300 python tools/probes/gpu_stage_probe.py config2_ell1 5000 4096 > gpurun_out/c_c2.log 2>&1; tail -1 gpurun_out/c_c2.log
timeout 300 python tools/probes/gpu_latency_probe.py > gpurun_out/c_lat.log 2>&1; tail -8 gpurun_out/c_lat.log
timeout 120 python tools/probes/gpu_coissue_probe.py > gpurun_out/c_coissue.log 2>&1; cat gpurun_out/c_coissue.log
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -o gpurun_out/c_c3 -f python bench.py --profile --no-extras --no-cpu-baseline > gpurun_out/c_ncu_c3.log 2>&1; echo "ncu c3 rc $?"
