#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/probes/gpu_bench_vs_probe.py 2>&1 | tail -8
