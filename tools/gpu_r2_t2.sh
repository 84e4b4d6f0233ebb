#!/bin/bash
run() { tag=$1; cfg=$2; shift 2; env "$@" timeout 300 python tools/probes/gpu_stage_probe.py $cfg 2>&1 | tail -1 | cut -c1-110 | sed "s/^/[$tag] /"; }
for cfg in "config3_gp 10000 8192" "config5_array 100000 8192"; do
run default "$cfg" X=1
run u2 "$cfg" VELA_B200_JIT_UNROLL=2
run u2_mb4 "$cfg" VELA_B200_JIT_UNROLL=2 VELA_B200_JIT_MINBLOCKS=4
run mb4 "$cfg" VELA_B200_JIT_MINBLOCKS=4
run t256_mb2 "$cfg" VELA_B200_CHAIN_THREADS=256 VELA_B200_JIT_MINBLOCKS=2
run g16 "$cfg" VELA_B200_CHAIN_GROUP=16
run g64 "$cfg" VELA_B200_CHAIN_GROUP=64
done
