#!/bin/bash
# Round 2, GPU call R: ncu of the rank-300 panel Cholesky and of the config-5 column sums / chain.
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:chol_panel -c 1 -o gpurun_out/r_panel -f python tools/probes/gpu_stage_probe.py config5_array 20000 2048 > gpurun_out/r_ncu.log 2>&1; echo "ncu rc $?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"colsum|chain" -c 3 -o gpurun_out/r_c5 -f python tools/probes/gpu_stage_probe.py config5_array 100000 2048 > gpurun_out/r_ncu_c5.log 2>&1; echo "ncu c5 rc $?"
