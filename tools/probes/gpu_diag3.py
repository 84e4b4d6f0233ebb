import sys, os
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle'); sys.path.insert(0, 'tests')
from __graft_entry__ import load_package
vb = load_package()
import oracle as O
name = sys.argv[1] if len(sys.argv) > 1 else "extra_dmjump_bits_wb"
ds = vb.synth.make_dataset(name, lambda m, t: vb.Pulsar(m, t))
psr = vb.Pulsar(ds.model, ds.toas); md = vb.ModelDescriptor(ds.model, ds.toas)
W = ds.walkers(64, seed=21)
got = psr.lnlike_vectorized(W); want = O.lnlike_batch(md, W)
rel = np.abs(got - want) / np.abs(want)
print('max rel', rel.max(), 'argmax', rel.argmax(), 'median', np.median(rel))
i = rel.argmax()
S, u = psr.sigma_and_u(W[i]); So, uo = O.sigma(md, W[i])
d = np.sqrt(np.diag(So)); Sn = So / np.outer(d, d)
print('cond of scaled Sigma', np.linalg.cond(Sn), 'Sigma rel err', np.max(np.abs(S - So) / np.outer(d, d)), 'u rel', np.max(np.abs(u - uo)) / np.abs(uo).max())
y, w = psr.resids_and_Ninvdiag(W[i]); yo, wo = O.residuals(md, W[i])
print('dy', np.max(np.abs(y - yo)[:len(ds.toas)]), 'dw', np.max(np.abs(w / wo - 1)), 'single lnlike gpu', psr.lnlike(W[i]), 'oracle', want[i], 'batch', got[i])
for B in (1, 8, 64):
    g = psr.lnlike_vectorized(W[i:i+1].repeat(B, 0))
    print('B', B, g[0], (g[0]-want[i])/abs(want[i]))
