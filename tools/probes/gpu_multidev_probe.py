"""In-process multi-device handle (what the Julia host uses): wall clock per host call, 8192 walkers per device."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import torch
from __graft_entry__ import load_package
vb = load_package()
nd = torch.cuda.device_count()
ds = vb.synth.make_dataset("config3_gp", lambda m, t: vb.Pulsar(m, t, devices=[0]))
for devs in ([0], list(range(nd))):
    psr = vb.Pulsar(ds.model, ds.toas, devices=devs)
    W = ds.walkers(8192 * len(devs), seed=1)
    for _ in range(3):
        out = psr.lnlike_vectorized(W)
    ts = []
    for _ in range(20):
        t0 = time.perf_counter(); out = psr.lnlike_vectorized(W); ts.append(time.perf_counter() - t0)
    print(f"devices {devs}: {len(W)} walkers, {min(ts) * 1e3:.3f} ms per call (median {np.median(ts) * 1e3:.3f}) -> "
          f"{len(W) / min(ts) / 1e6:.2f} M evals/s, finite {np.isfinite(out).mean():.3f}", flush=True)
