"""Probe: chain-kernel block size x resident blocks (VELA_B200_CHAIN_THREADS, VELA_B200_JIT_MINBLOCKS)."""
import os, sys, subprocess
exec(open('tools/probes/gpu_chain_probe.py').read().split("cases =")[0])   # writes gpurun_out/_cprobe.py
for name, ntoa, B in [("config3_gp", 10000, 8192), ("config4_dd_wb", 20000, 4096)]:
    for th, mb in [(256, 3), (128, 6), (128, 5), (128, 4), (192, 4), (192, 3), (64, 8), (64, 10)]:
        env = dict(os.environ, VELA_B200_CHAIN_THREADS=str(th), VELA_B200_JIT_MINBLOCKS=str(mb))
        r = subprocess.run([sys.executable, 'gpurun_out/_cprobe.py', name, str(ntoa), str(B)], env=env, capture_output=True, text=True, timeout=600)
        print(name, 'threads', th, 'minblocks', mb, r.stdout.strip() or r.stderr[-300:], flush=True)
