#!/bin/bash
# Round 2, GPU call J2: spin phase as a difference from the default parameters (kFlagDelta): full GPU suite, all configs with
# the route on and off, small batches.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=15 > gpurun_out/j2_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/j2_tests.log
tail -4 gpurun_out/j2_tests.log
run() { tag=$1; cfg=$2; shift 2; env "$@" timeout 300 python tools/probes/gpu_stage_probe.py $cfg 2>&1 | tail -1 | cut -c1-125 | sed "s/^/[$tag] /"; }
{
for cfg in "config2_ell1 5000 4096" "config3_gp 10000 8192" "config4_dd_wb 20000 8192" "config5_array 100000 8192"; do
run dd "$cfg" VELA_B200_DELTA=0
run delta "$cfg" X=1
run delta_mb6 "$cfg" VELA_B200_JIT_MINBLOCKS=6
done
} > gpurun_out/j2_stages.log 2>&1
cat gpurun_out/j2_stages.log
timeout 600 python tools/probes/gpu_latency_probe.py 2>&1 | head -4
