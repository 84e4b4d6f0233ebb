#!/bin/bash
# Round 2: the driver's multi-GPU launch at N = 8 (headline + extras: config 5 at 8192 walkers/rank = 65536 over 8 GPUs,
# strong scaling, in-process 8-device handle), then the reference arm.
mkdir -p gpurun_out
nvidia-smi -L | head -8 > gpurun_out/n8_gpus.txt
SECONDS=0
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/n8_bench.json 2> gpurun_out/n8_bench.err; echo "bench N=8 rc $? in $SECONDS s"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/n8_bench.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','n_gpus')}, l['e2e']['value'])
for k in ('other_configs','strong','inprocess'):
    print(k, json.dumps(l.get(k))[:1500])
PY
timeout 600 python -m pytest tests/test_gpu_multidevice.py -m gpu -q > gpurun_out/n8_multidev.log 2>&1; tail -3 gpurun_out/n8_multidev.log
