#!/bin/bash
for ks in default 4 8 16 32; do
  if [ $ks = default ]; then env X=1 timeout 300 python tools/probes/gpu_latency_probe.py 2>&1 | head -3 | sed "s/^/[ks $ks] /"
  else VELA_B200_SIGMA_KSPLIT=$ks timeout 300 python tools/probes/gpu_latency_probe.py 2>&1 | head -3 | sed "s/^/[ks $ks] /"; fi
done
