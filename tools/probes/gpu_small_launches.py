"""One small-batch call per batch size (for an ncu launch list): config3_gp at B = 16, 64, 256 after warm-up."""
import sys
import numpy as np
sys.path.insert(0, '.')
from __graft_entry__ import load_package
import torch
vb = load_package()
ds = vb.synth.make_dataset("config3_gp", lambda m, t: vb.Pulsar(m, t), ntoa=10000)
psr = vb.Pulsar(ds.model, ds.toas)
Ws = {B: ds.walkers(B, seed=1) for B in (16, 64, 256)}
for B, W in Ws.items():
    for _ in range(5):
        psr.lnlike_vectorized(W)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for B, W in Ws.items():
    psr.lnlike_vectorized(W)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
