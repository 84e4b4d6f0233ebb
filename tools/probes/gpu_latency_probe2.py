"""Wall-clock latency and stage times of one host call of config 3 at the batch sizes given on the command line."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from __graft_entry__ import load_package
vb = load_package()
ds = vb.synth.make_dataset("config3_gp", lambda m, t: vb.Pulsar(m, t), ntoa=10000)
psr = vb.Pulsar(ds.model, ds.toas)
for B in [int(a) for a in sys.argv[1:]]:
    W = ds.walkers(B, seed=1)
    for _ in range(20):
        psr.lnlike_vectorized(W)
    t0 = time.perf_counter()
    for _ in range(200):
        psr.lnlike_vectorized(W)
    dt = (time.perf_counter() - t0) / 200
    psr.set_stage_timing(True); psr.lnlike_vectorized(W); st = psr.last_stage_ms(); psr.set_stage_timing(False)
    print(f"B={B}: {dt*1e3:.3f} ms per call, kernels {st[4]:.3f} = prologue {st[0]:.3f} + chain {st[1]:.3f} + sigma {st[2]:.3f} + chol {st[3]:.3f}")
