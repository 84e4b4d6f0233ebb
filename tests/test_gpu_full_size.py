"""Full BASELINE.json sizes on the GPU: size-independent properties + an oracle check on a walker subset."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,nsub", [("config2_ell1", 64), ("config3_gp", 64), ("config4_dd_wb", 32)])
def test_full_size_config(name, nsub, vb, oracle):
    cfg = vb.synth.CONFIGS[name]
    ds = vb.synth.make_dataset(name, lambda m, t: vb.Pulsar(m, t))
    assert len(ds.toas) == cfg["ntoa"]
    psr = vb.Pulsar(ds.model, ds.toas)
    B = cfg["walkers"]
    W = ds.walkers(B, seed=5)
    out = psr.lnlike_vectorized(W)
    assert out.shape == (B,) and np.all(np.isfinite(out))
    # property: row independence — any sub-batch reproduces the same numbers (to rounding: a smaller batch may use
    # a different split-K factor in the Sigma contraction), and a permuted batch of the same size bit for bit
    idx = np.random.default_rng(1).choice(B, size=100, replace=False)
    np.testing.assert_allclose(psr.lnlike_vectorized(W[idx]), out[idx], rtol=1e-12)
    perm = np.random.default_rng(2).permutation(B)
    np.testing.assert_array_equal(psr.lnlike_vectorized(W[perm]), out[perm])
    # property: lnpost = lnprior + lnlike with a flat prior
    np.testing.assert_array_equal(psr._batch(psr._lib.vela_b200_lnpost_batch, W[:256], None), psr.lnlike_vectorized(W[:256]))
    # oracle on a subset
    md = vb.ModelDescriptor(ds.model, ds.toas)
    np.testing.assert_allclose(out[:nsub], oracle.lnlike_batch(md, W[:nsub]), rtol=1e-9)
    y, _ = psr.resids_and_Ninvdiag(ds.truth)
    yo, _ = oracle.residuals(md, ds.truth)
    assert np.max(np.abs(y - yo)[: len(ds.toas)]) < 1e-9


def test_config5_full_toa_count(vb, oracle):
    """BASELINE configs[4] at its full 100k TOAs and rank 300 (red + DM + chromatic GP); the walker count is cut to
    512 (the full 65 536 are 8 GPUs' worth) and the oracle checks four of them."""
    cfg = vb.synth.CONFIGS["config5_array"]
    ds = vb.synth.make_dataset("config5_array", lambda m, t: vb.Pulsar(m, t))
    assert len(ds.toas) == cfg["ntoa"] == 100000 and ds.info["rank"] == 300
    psr = vb.Pulsar(ds.model, ds.toas)
    assert psr.sigma_structured() and psr.chain_specialized()
    W = ds.walkers(512, seed=5)
    out = psr.lnlike_vectorized(W)
    assert np.all(np.isfinite(out))
    perm = np.random.default_rng(2).permutation(len(W))
    np.testing.assert_array_equal(psr.lnlike_vectorized(W[perm]), out[perm])
    md = vb.ModelDescriptor(ds.model, ds.toas)
    np.testing.assert_allclose(out[:4], oracle.lnlike_batch(md, W[:4]), rtol=1e-9)
