#!/bin/bash
# Round 2, GPU call A: parity tests with the new chain / Cholesky kernels, variant probes, the extended bench, ncu counts.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/a_gpu.txt
timeout 900 python -m pytest tests -m gpu -q --maxfail=15 > gpurun_out/a_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/a_tests.log
tail -5 gpurun_out/a_tests.log
timeout 600 python tools/probes/gpu_variants_probe.py config3_gp 10000 8192 > gpurun_out/a_var_c3.log 2>&1
cat gpurun_out/a_var_c3.log
timeout 600 python tools/probes/gpu_variants_probe.py config4_dd_wb 20000 8192 default,unroll2_mb4,mb4,chol_cta > gpurun_out/a_var_c4.log 2>&1
cat gpurun_out/a_var_c4.log
timeout 300 python tools/probes/gpu_stage_probe.py config5_array 100000 8192 > gpurun_out/a_c5.log 2>&1; tail -1 gpurun_out/a_c5.log
timeout 300 python tools/probes/gpu_stage_probe.py config2_ell1 5000 4096 > gpurun_out/a_c2.log 2>&1; tail -1 gpurun_out/a_c2.log
timeout 300 python tools/probes/gpu_latency_probe.py > gpurun_out/a_lat.log 2>&1; tail -12 gpurun_out/a_lat.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc $?"; tail -3 gpurun_out/a_bench.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.sum,smsp__inst_executed_op_shared_ld.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/a_ncu_metrics.csv python bench.py --profile --no-extras --no-cpu-baseline > gpurun_out/a_ncu.log 2>&1; echo "ncu rc $?"
