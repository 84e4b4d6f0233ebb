"""Chain-kernel time of one synthetic config: the same walker batch repeated vs batches alternating call by call."""
import sys
import numpy as np
sys.path.insert(0, '.')
from __graft_entry__ import load_package
vb = load_package()
name, ntoa, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
ds = vb.synth.make_dataset(name, lambda m, t: vb.Pulsar(m, t), ntoa=ntoa)
psr = vb.Pulsar(ds.model, ds.toas)
psr.set_stage_timing(True)
Ws = [ds.walkers(B, seed=s) for s in (1, 2, 1000, 1001)]
for _ in range(3):
    psr.lnlike_vectorized(Ws[0])
same = []
for _ in range(8):
    psr.lnlike_vectorized(Ws[0]); same.append(psr.last_stage_ms()[1])
alt = []
for i in range(8):
    psr.lnlike_vectorized(Ws[i % 4]); alt.append(psr.last_stage_ms()[1])
same2 = []
for _ in range(4):
    psr.lnlike_vectorized(Ws[3]); same2.append(psr.last_stage_ms()[1])
# identical values in a different host array
W0b = Ws[0].copy()
cp = []
for i in range(4):
    psr.lnlike_vectorized(W0b if i % 2 else Ws[0]); cp.append(psr.last_stage_ms()[1])
print("same batch     :", " ".join(f"{t:.3f}" for t in same))
print("alternating    :", " ".join(f"{t:.3f}" for t in alt))
print("same (last one):", " ".join(f"{t:.3f}" for t in same2))
print("copies of one  :", " ".join(f"{t:.3f}" for t in cp))
