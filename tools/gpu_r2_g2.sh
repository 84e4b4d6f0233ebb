#!/bin/bash
# Round 2, GPU call G2: ncu --set full of the config-5 step (100k TOAs, rank 300; 2048 walkers to keep the replays short).
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"vela_chain_spec|colsum_kernel|chol_panel" --launch-skip 8 --launch-count 4 -o gpurun_out/g2_c5 -f python tools/probes/gpu_stage_probe.py config5_array 100000 2048 > gpurun_out/g2_ncu.log 2>&1; echo "ncu rc $?"; tail -2 gpurun_out/g2_ncu.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"vela_chain_spec|colsum_kernel|chol_dmma" --launch-skip 8 --launch-count 4 -o gpurun_out/g2_c4 -f python tools/probes/gpu_stage_probe.py config4_dd_wb 20000 4096 > gpurun_out/g2_ncu4.log 2>&1; echo "ncu rc $?"; tail -2 gpurun_out/g2_ncu4.log
