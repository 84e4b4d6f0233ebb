#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/probes/gpu_chain_seed_probe.py config3_gp 10000 8192 2>&1 | tail -4
VELA_B200_DELTA=0 timeout 300 python tools/probes/gpu_chain_seed_probe.py config3_gp 10000 8192 2>&1 | tail -4 | sed 's/^/[dd] /'
