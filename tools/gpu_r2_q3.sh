#!/bin/bash
for a in plain lazyenv import_only torch_init_noalloc torch_after eagerenv_torch; do timeout 300 python tools/probes/gpu_bench_vs_probe3.py $a 2>&1 | tail -1; done
