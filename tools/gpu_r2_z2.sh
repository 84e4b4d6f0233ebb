#!/bin/bash
run() { tag=$1; cfg=$2; shift 2; env "$@" timeout 300 python tools/probes/gpu_stage_probe.py $cfg 2>&1 | tail -1 | cut -c1-110 | sed "s/^/[$tag] /"; }
for cfg in "config3_gp 10000 8192" "config5_array 100000 8192" "config2_ell1 5000 4096"; do
run default "$cfg" X=1
run sunregs "$cfg" VELA_B200_JIT_SUNREGS=1
done
