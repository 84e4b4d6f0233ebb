"""Why does the chain stage of config 3 take longer inside bench.py than in the stage probe?  Same handle, four ways."""
import sys
import numpy as np
sys.path.insert(0, '.')
import torch
from __graft_entry__ import load_package
import bench
vb = load_package()
B = 8192
ds = vb.synth.make_dataset("config3_gp", lambda m, t: vb.Pulsar(m, t, devices=[0]))
psr = vb.Pulsar(ds.model, ds.toas, devices=[0])
Ws = [ds.walkers(B, seed=1000 + i) for i in range(4)]
psr.set_stage_timing(True)
def chain_host(n=4):
    out = []
    for i in range(n):
        psr.lnlike_vectorized(Ws[i % 4]); out.append(psr.last_stage_ms()[1])
    return " ".join(f"{t:.3f}" for t in out)
print("host path, no priors      :", chain_host())
dev = [torch.from_numpy(w).cuda() for w in Ws]
d_out = torch.empty(B, dtype=torch.float64, device="cuda")
zero_lp = torch.zeros(B, dtype=torch.float64, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
def chain_dev(lp, n=4):
    out = []
    for i in range(n):
        psr.lnpost_device(dev[i % 4].data_ptr(), B, lp[i % 4].data_ptr() if isinstance(lp, list) else lp.data_ptr(), d_out.data_ptr(), stream)
        torch.cuda.synchronize(); out.append(psr.last_stage_ms()[1])
    return " ".join(f"{t:.3f}" for t in out)
print("device path, zero lnprior :", chain_dev(zero_lp))
pri = bench.make_priors(vb, ds)
psr.set_priors(pri)
lnpr = [torch.from_numpy(np.ascontiguousarray(vb.priors.lnprior(pri, w))).cuda() for w in Ws]
print("finite priors             :", [float(torch.isfinite(l).double().mean()) for l in lnpr])
print("device path, bench priors :", chain_dev(lnpr))
print("host path after priors    :", chain_host())
s2 = torch.cuda.Stream()
with torch.cuda.stream(s2):
    stream = s2.cuda_stream
    print("device path, side stream  :", chain_dev(lnpr))
