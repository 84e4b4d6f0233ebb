"""Perf probe: resident blocks per SM of the NVRTC chain kernel (VELA_B200_JIT_MINBLOCKS)."""
import os, sys, subprocess
code = r'''
import sys, numpy as np
sys.path.insert(0, '.')
from __graft_entry__ import load_package
vb = load_package()
name, ntoa, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
ds = vb.synth.make_dataset(name, lambda m, t: vb.Pulsar(m, t), ntoa=ntoa)
psr = vb.Pulsar(ds.model, ds.toas)
W = ds.walkers(B, seed=1)
psr.set_stage_timing(True)
psr.lnlike_vectorized(W)
best = 1e9
for _ in range(4):
    out = psr.lnlike_vectorized(W); st = psr.last_stage_ms(); best = min(best, st[1])
print(f"chain {best:.3f} ms spec {psr.chain_specialized()} sum {out.sum():.10e}")
'''
os.makedirs('gpurun_out', exist_ok=True)
open('gpurun_out/_cprobe.py', 'w').write(code)
cases = [("config3_gp", 10000, 8192), ("config2_ell1", 5000, 4096), ("config4_dd_wb", 20000, 4096), ("extra_ddk_wb", 5000, 2048), ("extra_ell1h_fd_wavex", 5000, 2048)]
for name, ntoa, B in cases:
    for mb in (2, 3, 4):
        env = dict(os.environ, VELA_B200_JIT_MINBLOCKS=str(mb))
        r = subprocess.run([sys.executable, 'gpurun_out/_cprobe.py', name, str(ntoa), str(B)], env=env, capture_output=True, text=True, timeout=600)
        print(name, ntoa, B, 'minblocks', mb, r.stdout.strip() or r.stderr[-300:], flush=True)
