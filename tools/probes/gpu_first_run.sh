#!/bin/bash
# first GPU bring-up: parity tests + FP64 peak micro-benchmarks
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -30
python - <<'PY'
import sys; sys.path.insert(0,'.')
from __graft_entry__ import load_package
vb=load_package()
L=vb.lib.lib()
for rep in range(2):
    print('DFMA peak TFLOP/s', L.vela_b200_fp64_peak_tflops(0,0), 'DMMA peak TFLOP/s', L.vela_b200_fp64_peak_tflops(0,1))
PY
