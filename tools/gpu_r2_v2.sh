#!/bin/bash
for sh in default 0 1 2 4; do
  if [ $sh = default ]; then timeout 300 python tools/probes/gpu_latency_probe2.py 16 32 64 128 256 2>&1 | sed "s/^/[shape $sh] /"
  else VELA_B200_COLSUM_SHAPE=$sh timeout 300 python tools/probes/gpu_latency_probe2.py 16 32 64 128 256 2>&1 | sed "s/^/[shape $sh] /"; fi
done
