"""Small end-to-end run that touches every kernel variant (white, ECORR, Woodbury in shared and global scratch, structured
and dense Sigma paths, conditional mean) — a quick sanity pass after kernel changes."""
import os, sys
import numpy as np
sys.path.insert(0, '.')
from __graft_entry__ import load_package
vb = load_package()
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for name in ("NGC6440E", "sim3_gp", "sim_dmgp_wb", "sim2", "sim_sw_wb"):          # white, rank 180 (shared), rank 248 (global), ECORR, wideband
    model, toas, _ = vb.readwrite.load_pulsar_npz(os.path.join(root, "tests", "golden", name + ".npz"))
    z = np.load(os.path.join(root, "tests", "golden", name + ".npz"))
    psr = vb.Pulsar(model, toas)
    W = z["golden_walkers"][:9]
    out = psr.lnpost_vectorized(W)
    print(name, "structured", psr.sigma_structured(), "jit", psr.chain_specialized(), out[:2])
    if psr.has_marginalized_gp_noise:
        psr.marginalized_offset_mean_vectorized(W[:3], with_cf=True)
        psr.sigma_and_u(W[0])
for cfg, ntoa in (("extra_ecorr_gp", None), ("config4_dd_wb", 300), ("extra_free_gp_wb", None)):
    ds = vb.synth.make_dataset(cfg, lambda m, t: vb.Pulsar(m, t), ntoa=ntoa)
    psr = vb.Pulsar(ds.model, ds.toas)
    print(cfg, psr.lnlike_vectorized(ds.walkers(70, seed=1))[:2])
    os.environ["VELA_B200_STRUCTURED"] = "0"
    dense = vb.Pulsar(ds.model, ds.toas)
    del os.environ["VELA_B200_STRUCTURED"]
    if dense.has_marginalized_gp_noise:
        print(cfg, "dense", dense.lnlike_vectorized(ds.walkers(70, seed=1))[:2])
print("done")
