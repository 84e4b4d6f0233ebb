"""Chain-kernel time of one synthetic config for several walker batches (seeds): does it depend on the batch?"""
import sys
import numpy as np
sys.path.insert(0, '.')
from __graft_entry__ import load_package
vb = load_package()
name, ntoa, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
ds = vb.synth.make_dataset(name, lambda m, t: vb.Pulsar(m, t), ntoa=ntoa)
psr = vb.Pulsar(ds.model, ds.toas)
psr.set_stage_timing(True)
for seed in (1, 2, 1000, 1001, 1002, 1003, 1, 1000):
    W = ds.walkers(B, seed=seed)
    ts = []
    for _ in range(3):
        out = psr.lnlike_vectorized(W)
        ts.append(psr.last_stage_ms()[1])
    print(f"seed {seed}: chain {ts[0]:.3f} {ts[1]:.3f} {ts[2]:.3f} ms, finite {np.isfinite(out).mean():.3f}")
