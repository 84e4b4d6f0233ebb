"""Small-batch (emcee-sized) calls: wall clock per host call against the device stage times."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from __graft_entry__ import load_package
vb = load_package()
name = sys.argv[1] if len(sys.argv) > 1 else "config3_gp"
ds = vb.synth.make_dataset(name, lambda m, t: vb.Pulsar(m, t))
psr = vb.Pulsar(ds.model, ds.toas)
W = ds.walkers(1024, seed=1)
for B in (1, 16, 64, 256, 1024):
    psr.set_stage_timing(False)
    for _ in range(5):
        psr.lnlike_vectorized(W[:B])
    ts = []
    for _ in range(50):
        t0 = time.perf_counter(); psr.lnlike_vectorized(W[:B]); ts.append(time.perf_counter() - t0)
    psr.set_stage_timing(True)
    psr.lnlike_vectorized(W[:B]); psr.lnlike_vectorized(W[:B])
    st = psr.last_stage_ms()
    print(f"{name} B={B}: wall min {min(ts) * 1e3:.3f} median {np.median(ts) * 1e3:.3f} ms | device: prologue {st[0]:.3f} chain {st[1]:.3f} "
          f"sigma {st[2]:.3f} chol {st[3]:.3f} sum {st[4]:.3f} ms", flush=True)
