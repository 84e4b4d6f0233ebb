#!/bin/bash
for a in "notorch nodev" "notorch dev" "torch nodev" "torch dev"; do timeout 300 python tools/probes/gpu_bench_vs_probe2.py $a 2>&1 | tail -1; done
