"""Device-side priors (lnprior / prior_transform) against the host-side mirror of Distributions.jl."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _all_families(vb, nd):
    P = vb.priors
    fams = [P.Uniform(-1.0, 3.0), P.Normal(0.5, 2.0), P.LogNormal(0.0, 0.25), P.LogUniform(1e-3, 10.0),
            P.KINPriorDistribution(), P.SINIPriorDistribution(), P.STIGMAPriorDistribution(),
            P.SHAPMAXPriorDistribution(), P.PXPriorDistribution(2.0), P.LatitudePriorDistribution(),
            P.TruncatedNormal(0.3, 1.5, -1.0, 2.0)]
    return [P.SimplePrior(f"p{k}", fams[k % len(fams)]) for k in range(nd)]


def test_device_priors_match_host(vb, oracle):
    ds = vb.synth.make_dataset("extra_ell1_fbx", lambda m, t: vb.Pulsar(m, t), ntoa=300)
    psr = vb.Pulsar(ds.model, ds.toas)
    nd = psr.ndim
    priors = _all_families(vb, nd)
    assert psr.set_priors(priors)
    rng = np.random.default_rng(4)
    q = rng.uniform(0.01, 0.99, size=(257, nd))
    x_host = vb.priors.prior_transform(priors, q)
    x_dev = psr.prior_transform_vectorized(q)
    np.testing.assert_allclose(x_dev, x_host, rtol=1e-12, atol=1e-14)
    lp_host = vb.priors.lnprior(priors, x_host)
    np.testing.assert_allclose(psr.lnprior_vectorized(x_host), lp_host, rtol=1e-12)
    # outside the support: -Inf on both sides, and lnpost returns the prior itself (posterior.jl:9-12)
    bad = x_host.copy()
    bad[5, 0] = 100.0          # Uniform(-1, 3)
    lp = psr.lnprior_vectorized(bad)
    assert lp[5] == -np.inf and np.isfinite(lp[4])


def test_lnpost_uses_device_priors(vb, golden, oracle):
    g = golden("NGC6440E")
    psr = vb.Pulsar(g.model, g.toas)
    assert psr.device_priors                      # Uniform / LogNormal / LogUniform all have device twins
    W = g.walkers[:32].copy()
    W[7, 0] += 1.0                                # outside the RAJ box
    post = psr.lnpost_vectorized(W)
    lp = vb.priors.lnprior(g.model.priors, W)
    ll = oracle.lnlike_batch(g.md, W)
    want = np.where(np.isfinite(lp), lp + ll, lp)
    assert post[7] == -np.inf
    ok = np.isfinite(want)
    np.testing.assert_allclose(post[ok], want[ok], rtol=1e-9)


def test_reference_prior_literals_on_device(vb, golden):
    """test/test_priors.jl:3-44 through the device twins: PHOFF, F0 (truncated Normal, lower = 0), F1."""
    P = vb.priors
    g = golden("NGC6440E")
    psr = vb.Pulsar(g.model, g.toas)
    nd = psr.ndim
    priors = [P.SimplePrior("PHOFF", P.Uniform(-0.5, 0.5)), P.SimplePriorMulti("F", 1, P.TruncatedNormal(100.0, 1e-7, 0.0, np.inf)),
              P.SimplePriorMulti("F", 2, P.Uniform(-1.01e-14, -0.9e-14))]
    priors += [P.SimplePrior(f"x{k}", P.Uniform(-1e9, 1e9)) for k in range(nd - 3)]
    assert psr.set_priors(priors)
    pad = np.zeros(nd - 3)
    good = np.concatenate([[0.1, 100.0, -1e-14], pad])
    bad = np.concatenate([[1.1, -200.0, -2e-14], pad])
    lp = psr.lnprior_vectorized(np.stack([good, bad]))
    assert np.isfinite(lp[0]) and lp[1] == -np.inf
    np.testing.assert_allclose(lp[0], P.lnprior(priors, good), rtol=1e-12)
    cube = np.concatenate([[0.5, 0.5, 0.0], np.full(nd - 3, 0.5)])
    got = psr.prior_transform_vectorized(cube[None, :])[0]
    np.testing.assert_allclose(got[:3], [0.0, 100.0, -1.01e-14], rtol=1e-12, atol=1e-30)
