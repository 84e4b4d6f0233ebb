"""Run one full-size batch of a synthetic config between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import sys
import numpy as np
sys.path.insert(0, '.')
import torch
from __graft_entry__ import load_package
vb = load_package()
name, ntoa, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
ds = vb.synth.make_dataset(name, lambda m, t: vb.Pulsar(m, t), ntoa=ntoa)
psr = vb.Pulsar(ds.model, ds.toas)
W = ds.walkers(B, seed=1)
psr.lnlike_vectorized(W)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
out = psr.lnlike_vectorized(W)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("finite", np.isfinite(out).mean())
