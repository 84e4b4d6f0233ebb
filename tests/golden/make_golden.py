"""Regenerates tests/golden/*.npz from the reference's JLSO test fixtures.

Run in the dev container (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
For each fixture it stores the (model, TOAs) pair in the portable npz container plus oracle outputs
at the default parameter values and on a small seeded walker batch, so that
  * `-m "not gpu"` tests can check the oracle against the recorded numbers (regression pin), and
  * `-m gpu` tests can check the CUDA path against the same numbers without /root/reference.
The oracle itself is pinned against the reference's own test bounds in tests/test_oracle_fixtures.py.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from __graft_entry__ import load_package  # noqa: E402

vb = load_package()
import oracle as O  # noqa: E402

REF = "/root/reference/test/datafiles"
FIXTURES = ["NGC6440E", "sim_sw.wb", "sim3.gp", "sim_dmgp_wb", "sim2"]
# walkers per fixture: BASELINE.json configs[0] names a batch of 1024 samples for NGC6440E; the others keep 64 (their
# oracle evaluations are the slow part of the CPU suite)
N_WALKERS = {"NGC6440E": 1024}


def walkers(model, n, seed):
    """Seeded walker batch: free parameters scattered inside their (cheat) priors around the defaults."""
    rng = np.random.default_rng(seed)
    x0 = model.read_param_values_to_vector()
    out = np.tile(x0, (n, 1))
    for k, pr in enumerate(model.priors):
        d = pr.distribution
        if type(d).__name__ == "Uniform":
            width = d.b - d.a
            out[:, k] = np.clip(x0[k] + 0.01 * width * rng.standard_normal(n), d.a, d.b)
        elif type(d).__name__ == "LogNormal":
            out[:, k] = x0[k] * np.exp(0.1 * rng.standard_normal(n))
        elif type(d).__name__ == "LogUniform":
            out[:, k] = np.exp(rng.uniform(np.log(d.a), np.log(d.b), n))
    out[0] = x0
    return out


def main():
    for name in FIXTURES:
        model, toas = vb.load_pulsar_data(os.path.join(REF, name + ".jlso"))
        md = vb.ModelDescriptor(model, toas)
        x0 = model.read_param_values_to_vector()
        y, w = O.residuals(md, x0)
        W = walkers(model, N_WALKERS.get(name, 64), seed=20260101)
        extra = {
            "source": f"test/datafiles/{name}.jlso",
            "lnlike_default": O.lnlike(md, x0),
            "chi2_default": O.chi2(md, x0) if md.desc.kernel_kind in (0, 1) else None,
        }
        out = os.path.join(ROOT, "tests", "golden", name.replace(".", "_") + ".npz")
        vb.readwrite.save_pulsar_npz(out, model, toas, extra)
        z = dict(np.load(out))
        z.update(golden_y=y, golden_w=w, golden_walkers=W, golden_lnlike=O.lnlike_batch(md, W))
        if md.n_basis:
            z["golden_phiinv"] = O.noise_weights_inv(md, x0)
        np.savez_compressed(out, **z)
        print(name, os.path.getsize(out), extra)


if __name__ == "__main__":
    main()
