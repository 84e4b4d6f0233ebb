#!/bin/bash
# Round 2, GPU call Y: chain kernel without in-line fallback branches (experiment), with and without 2-walker unrolling.
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python tools/probes/gpu_stage_probe.py config3_gp 10000 8192 2>&1 | tail -1 | sed "s/^/[$tag] /"; }
{
run default X=1
run nofb VELA_B200_JIT_EXPERIMENT=1 VELA_B200_JIT_DUMP=gpurun_out/y_nofb
run nofb_u2 VELA_B200_JIT_EXPERIMENT=1 VELA_B200_JIT_UNROLL=2
run nofb_u2_mb4 VELA_B200_JIT_EXPERIMENT=1 VELA_B200_JIT_UNROLL=2 VELA_B200_JIT_MINBLOCKS=4
run nofb_mb6 VELA_B200_JIT_EXPERIMENT=1 VELA_B200_JIT_MINBLOCKS=6
run default_dump VELA_B200_JIT_DUMP=gpurun_out/y_default
} > gpurun_out/y_stages.log 2>&1
cat gpurun_out/y_stages.log; ls gpurun_out | grep y_
