#!/bin/bash
# Round 2, GPU call Z: speculative straight-line chain (default) vs branching chain, all configs; full GPU suite.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=15 > gpurun_out/z_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/z_tests.log
tail -4 gpurun_out/z_tests.log
run() { tag=$1; cfg=$2; shift 2; env "$@" timeout 300 python tools/probes/gpu_stage_probe.py $cfg 2>&1 | tail -1 | cut -c1-150 | sed "s/^/[$tag] /"; }
{
for cfg in "config2_ell1 5000 4096" "config3_gp 10000 8192" "config4_dd_wb 20000 8192" "config5_array 100000 8192"; do
run branching "$cfg" VELA_B200_JIT_SPECULATE=0
run spec "$cfg" X=1
run spec_u2 "$cfg" VELA_B200_JIT_UNROLL=2
run spec_u2_mb4 "$cfg" VELA_B200_JIT_UNROLL=2 VELA_B200_JIT_MINBLOCKS=4
run spec_mb4 "$cfg" VELA_B200_JIT_MINBLOCKS=4
done
} > gpurun_out/z_stages.log 2>&1
cat gpurun_out/z_stages.log
VELA_B200_JIT_DUMP=gpurun_out/z_c3.cubin timeout 300 python tools/probes/gpu_stage_probe.py config3_gp 10000 1024 > /dev/null 2>&1
