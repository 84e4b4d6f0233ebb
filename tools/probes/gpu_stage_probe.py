"""Stage times (prologue, chain, Sigma, Cholesky/finish; ms) of one synthetic config at a given size: name ntoa walkers."""
import sys
import numpy as np
sys.path.insert(0, '.')
from __graft_entry__ import load_package
vb = load_package()
name, ntoa, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
ds = vb.synth.make_dataset(name, lambda m, t: vb.Pulsar(m, t), ntoa=ntoa)
psr = vb.Pulsar(ds.model, ds.toas)
W = ds.walkers(B, seed=1)
psr.set_stage_timing(True)
for _ in range(3):
    out = psr.lnlike_vectorized(W)
st = psr.last_stage_ms()
print(f"{name} ntoa={ntoa} B={B} rank={ds.info['rank']}: prologue {st[0]:.3f} chain {st[1]:.3f} sigma {st[2]:.3f} chol {st[3]:.3f} "
      f"total {st[4]:.3f} ms, finite {np.isfinite(out).mean():.3f}", end="")
if ds.info['rank']:
    si = psr.sigma_info(B)
    print(f" | sigma: two_stage {si['two_stage']} cols {si['table_columns']} s1 {si['stage1_columns']} s2 {si['stage2_columns']} "
          f"k2 {si['moments_per_block']} cells {si['cells']} -> {si['executed_flops_per_walker'] * B / (st[2] * 1e-3) / 1e12:.1f} TFLOP/s executed")
else:
    print()
