#!/bin/bash
# Round 2, GPU call B: full ncu captures (source-level) of the config-3 step and of the config-4 chain kernel.
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -o gpurun_out/b_c3 -f python bench.py --profile --no-extras --no-cpu-baseline > gpurun_out/b_ncu_c3.log 2>&1; echo "ncu c3 rc $?"
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:vela_chain_spec -o gpurun_out/b_c4 -f python bench.py --profile --profile-config config4_dd_wb --no-extras --no-cpu-baseline > gpurun_out/b_ncu_c4.log 2>&1; echo "ncu c4 rc $?"
ls -la gpurun_out/b_*
