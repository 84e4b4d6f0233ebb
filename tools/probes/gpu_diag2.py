import sys, os
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle'); sys.path.insert(0, 'tests')
from __graft_entry__ import load_package
vb = load_package()
import oracle as O
ds = vb.synth.make_dataset("config3_gp", lambda m, t: vb.Pulsar(m, t), ntoa=1200)
psr = vb.Pulsar(ds.model, ds.toas); md = vb.ModelDescriptor(ds.model, ds.toas)
So, uo = O.sigma(md, ds.truth)
def chk(tag):
    S, u = psr.sigma_and_u(ds.truth)
    print(tag, 'S00', S[0,0], So[0,0], 'max rel', np.max(np.abs(S-So))/np.abs(So).max())
chk('fresh')
y, w = psr.resids_and_Ninvdiag(ds.truth)
chk('after resids')
W = ds.walkers(96, seed=21)
psr.lnlike_vectorized(W)
chk('after batch 96')
chk('again')
psr.lnlike_vectorized(W[:5])
chk('after batch 5')
