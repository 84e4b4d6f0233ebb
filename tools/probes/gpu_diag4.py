import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests"); sys.path.insert(0, "oracle")
from __graft_entry__ import load_package
import numpy as np, os
vb = load_package()
import oracle as O
nm = sys.argv[1] if len(sys.argv) > 1 else "sim_dmgp_wb"
m, t, ex = vb.readwrite.load_pulsar_npz(f"tests/golden/{nm}.npz")
z = np.load(f"tests/golden/{nm}.npz")
W = z["golden_walkers"]; want = z["golden_lnlike"]
psr = vb.Pulsar(m, t)
got = psr.lnlike_vectorized(W)
rel = np.abs(got - want) / np.abs(want)
print("B", len(W), "max rel", rel.max(), "argmax", rel.argmax(), got[:4], want[:4])
for B in (1, 2, 7, 64, len(W)):
    g = psr.lnlike_vectorized(W[:B]); print(B, np.max(np.abs(g - want[:B]) / np.abs(want[:B])))
