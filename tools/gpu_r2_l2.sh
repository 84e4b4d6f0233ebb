#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; cfg=$2; shift 2; env "$@" timeout 300 python tools/probes/gpu_stage_probe.py $cfg 2>&1 | tail -1 | cut -c1-125 | sed "s/^/[$tag] /"; }
{
cfg="config3_gp 10000 8192"
run base "$cfg" VELA_B200_LIB=$PWD/build/ab/libvela_b200_base.so
run delta "$cfg" VELA_B200_LIB=$PWD/build/ab/libvela_b200_delta.so
run delta_off "$cfg" VELA_B200_LIB=$PWD/build/ab/libvela_b200_delta.so VELA_B200_DELTA=0
run base "$cfg" VELA_B200_LIB=$PWD/build/ab/libvela_b200_base.so
} > gpurun_out/l2_stages.log 2>&1
cat gpurun_out/l2_stages.log
