#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:vela_chain_spec --launch-skip 2 --launch-count 1 -o gpurun_out/o2_probe -f python tools/probes/gpu_stage_probe.py config3_gp 10000 8192 > gpurun_out/o2_ncu1.log 2>&1; echo "rc $?"
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:vela_chain_spec -o gpurun_out/o2_bench -f python bench.py --profile --no-extras --no-cpu-baseline > gpurun_out/o2_ncu2.log 2>&1; echo "rc $?"
