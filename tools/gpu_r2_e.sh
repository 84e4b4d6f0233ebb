#!/bin/bash
# Round 2, GPU call E (2 GPUs): the bench's multi-rank blocks (strong scaling, in-process handle, config 5 per rank) and the
# two-device handle test.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/e_gpus.txt
timeout 600 python -m pytest tests/test_gpu_multidevice.py tests/test_gpu_synth.py -m gpu -q > gpurun_out/e_tests.log 2>&1; tail -3 gpurun_out/e_tests.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/e_bench2.json 2> gpurun_out/e_bench2.err; echo "bench2 rc $?"; tail -3 gpurun_out/e_bench2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/e_ref2.json 2> gpurun_out/e_ref2.err; echo "ref2 rc $?"
