"""Chain stage of config 3: does importing torch / passing devices=[0] change it?  argv: torch|notorch  dev|nodev"""
import sys
import numpy as np
sys.path.insert(0, '.')
if sys.argv[1] == "torch":
    import torch
    torch.cuda.init(); torch.zeros(1, device="cuda")
from __graft_entry__ import load_package
vb = load_package()
kw = dict(devices=[0]) if sys.argv[2] == "dev" else {}
ds = vb.synth.make_dataset("config3_gp", lambda m, t: vb.Pulsar(m, t, **kw))
psr = vb.Pulsar(ds.model, ds.toas, **kw)
W = ds.walkers(8192, seed=1000)
psr.set_stage_timing(True)
out = []
for i in range(5):
    psr.lnlike_vectorized(W); out.append(psr.last_stage_ms()[1])
print(sys.argv[1:], " ".join(f"{t:.3f}" for t in out))
