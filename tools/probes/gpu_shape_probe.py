"""Perf probe: K3 launch shapes (VELA_B200_SIGMA_SHAPE index) on a small-rank and a large-rank problem."""
import os, sys, subprocess, json
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
for _ in range(3):
    out = psr.lnlike_vectorized(W); best = min(best, psr.last_stage_ms()[2])
p = ds.info["rank"]; nr = ntoa
print(f"sigma {best:.3f} ms  {(nr*p*(p+1)+3*nr*p)*B/(best*1e-3)/1e12:.2f} TF  finite {np.isfinite(out).mean():.2f}")
'''
os.makedirs('gpurun_out', exist_ok=True)
open('gpurun_out/_probe.py', 'w').write(code)
for name, ntoa, B in (("config3_gp", 10000, 8192), ("config5_array", 10000, 1024)):
    for shape in ([0, 1, 2, 3, 9, 10] if name == "config3_gp" else [5, 6, 7, 11]):
        env = dict(os.environ, VELA_B200_SIGMA_SHAPE=str(shape))
        r = subprocess.run([sys.executable, 'gpurun_out/_probe.py', name, str(ntoa), str(B)], env=env, capture_output=True, text=True, timeout=300)
        print(name, 'shape', shape, r.stdout.strip() or r.stderr[-300:])
