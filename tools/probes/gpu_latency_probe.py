"""Probe: wall-clock latency of one host call (numpy in, numpy out) at emcee-sized batches."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from __graft_entry__ import load_package
vb = load_package()
for name, ntoa in (("config3_gp", 10000), ("config2_ell1", 5000)):
    ds = vb.synth.make_dataset(name, lambda m, t: vb.Pulsar(m, t), ntoa=ntoa)
    psr = vb.Pulsar(ds.model, ds.toas)
    for B in (16, 64, 256, 1024):
        W = ds.walkers(B, seed=1)
        for _ in range(20):
            psr.lnlike_vectorized(W)
        t0 = time.perf_counter()
        n = 200
        for _ in range(n):
            psr.lnlike_vectorized(W)
        dt = (time.perf_counter() - t0) / n
        psr.set_stage_timing(True); psr.lnlike_vectorized(W); st = psr.last_stage_ms(); psr.set_stage_timing(False)
        print(f"{name} B={B}: {dt*1e3:.3f} ms per call ({B/dt:.0f} evals/s), kernels {st[4]:.3f} ms = prologue {st[0]:.3f} + chain {st[1]:.3f} + sigma {st[2]:.3f} + chol {st[3]:.3f}")
