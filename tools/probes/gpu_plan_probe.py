"""Perf probe: K3 plan knobs (VELA_B200_SIGMA_MT, VELA_B200_SIGMA_KSPLIT) at several problem sizes."""
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
best = 1e9; bc = 1e9; ba = 1e9
for _ in range(4):
    out = psr.lnlike_vectorized(W); st = psr.last_stage_ms(); best = min(best, st[2]); bc = min(bc, st[3]); ba = min(ba, st[4])
p = ds.info["rank"]; nr = ntoa * (2 if ds.info["wideband"] else 1)
print(f"sigma {best:.3f} ms chol {bc:.3f} all {ba:.3f}  {(nr*p*(p+1)+3*nr*p)*B/(best*1e-3)/1e12:.2f} TF  finite {np.isfinite(out).mean():.2f} sum {out.sum():.10e}")
'''
os.makedirs('gpurun_out', exist_ok=True)
open('gpurun_out/_probe.py', 'w').write(code)
cases = [("config3_gp", 10000, 8192), ("config3_gp", 10000, 64), ("config3_gp", 10000, 1024), ("config4_dd_wb", 20000, 4096), ("config5_array", 20000, 2048)]
plans = [(0, 0)] if len(sys.argv) > 1 and sys.argv[1] == "auto" else [(0, 0), (4, 1), (4, 2), (4, 3)]
for name, ntoa, B in cases:
    for mt, ks in plans:
        env = dict(os.environ)
        if mt:
            env["VELA_B200_SIGMA_MT"] = str(mt)
            env["VELA_B200_SIGMA_KSPLIT"] = str(ks)
        r = subprocess.run([sys.executable, 'gpurun_out/_probe.py', name, str(ntoa), str(B)], env=env, capture_output=True, text=True, timeout=600)
        print(name, ntoa, B, 'mt', mt, 'ks', ks, r.stdout.strip() or r.stderr[-300:], flush=True)
