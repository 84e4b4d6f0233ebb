"""Edge cases of the C-ABI entry points on the GPU: empty/ragged batches, strides, failure modes."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ngc(vb, golden):
    g = golden("NGC6440E")
    return g, vb.Pulsar(g.model, g.toas)


@pytest.fixture(scope="module")
def gp(vb, golden):
    g = golden("sim3_gp")
    return g, vb.Pulsar(g.model, g.toas)


def test_empty_batch(ngc):
    g, psr = ngc
    assert psr.lnlike_vectorized(np.zeros((0, len(g.x0)))).shape == (0,)


@pytest.mark.parametrize("B", [1, 2, 31, 33, 63, 64])
def test_ragged_batch_sizes(B, ngc, gp, oracle):
    for g, psr in (ngc, gp):
        W = np.resize(g.walkers, (B, g.walkers.shape[1]))
        np.testing.assert_allclose(psr.lnlike_vectorized(W), oracle.lnlike_batch(g.md, W), rtol=1e-9)


def test_batch_is_independent_of_order_and_neighbours(gp):
    """Each walker's result depends only on its own row (posterior.jl:21-23)."""
    g, psr = gp
    W = g.walkers[:40]
    a = psr.lnlike_vectorized(W)
    perm = np.random.default_rng(0).permutation(len(W))
    np.testing.assert_array_equal(psr.lnlike_vectorized(W[perm]), a[perm])
    np.testing.assert_array_equal(psr.lnlike_vectorized(W), a)            # deterministic: no atomics
    # a different batch SIZE may pick a different split-K factor for the Sigma contraction, i.e. a different
    # (still fixed) summation order over TOAs: equal to rounding, not bit for bit
    np.testing.assert_allclose(psr.lnlike_vectorized(W[:7]), a[:7], rtol=1e-13)
    sub = psr.lnlike_vectorized(W[:7])
    np.testing.assert_array_equal(psr.lnlike_vectorized(W[6::-1]), sub[::-1])


def test_column_major_and_strided_inputs(ngc):
    """A Julia Matrix is column-major, numpy through PythonCall row-major (posterior.jl:22 indexes either)."""
    g, psr = ngc
    W = g.walkers[:20]
    a = psr.lnlike_vectorized(W)
    np.testing.assert_array_equal(psr.lnlike_vectorized(np.asfortranarray(W)), a)
    big = np.zeros((40, 2 * W.shape[1]))
    big[::2, ::2] = W
    np.testing.assert_array_equal(psr.lnlike_vectorized(big[::2, ::2]), a)


def test_full_parameter_vector_is_accepted(ngc):
    """calc_lnlike takes the free vector or the full NamedTuple (wls_likelihood.jl:46-47)."""
    g, psr = ngc
    full = g.model.read_params(g.x0)
    assert psr.lnlike(full) == psr.lnlike(g.x0)


def test_wrong_number_of_free_parameters_is_an_error(vb, ngc):
    g, psr = ngc
    with pytest.raises(vb.VelaB200Error, match="free parameters"):     # parameter.jl:162
        psr.lnlike_vectorized(np.zeros((3, len(g.x0) + 1)))


def test_lnpost_prior_short_circuit(vb, ngc):
    """posterior.jl:9-12: a non-finite prior is returned as is; otherwise lnprior + lnlike."""
    g, psr = ngc
    W = g.walkers[:8].copy()
    W[3, 0] += 1.0                                                      # RAJ outside its prior box
    lp = vb.priors.lnprior(g.model.priors, W)
    post = psr.lnpost_vectorized(W)
    like = psr.lnlike_vectorized(W)
    assert post[3] == -np.inf and np.isfinite(like[3])
    ok = np.isfinite(lp)
    np.testing.assert_array_equal(post[ok], lp[ok] + like[ok])
    assert psr.lnpost(g.x0) == post[0] or np.isclose(psr.lnpost(g.x0), lp[0] + like[0], rtol=1e-15)


def test_not_positive_definite_gives_minus_inf(vb, gp):
    """gls_likelihood.jl:211-213.  A hugely negative log-amplitude makes Phi^-1 overflow -> non-finite pivot."""
    g, psr = gp
    names = g.model.get_free_param_names()
    W = g.walkers[:4].copy()
    W[1, names.index("TNREDAMP")] = 400.0        # 10^400 = Inf amplitude -> Phi^-1 = 0, log(0): -Inf / NaN path
    W[2, names.index("TNREDGAM")] = np.nan
    out = psr.lnlike_vectorized(W)
    assert np.isfinite(out[0]) and np.isfinite(out[3])
    assert out[2] == -np.inf and not np.isnan(out).any()


def test_invalid_eccentricity_gives_minus_inf(vb, oracle):
    """orbit.jl:26 asserts 0 < e < 1; across the ABI that sample is -Inf instead of an exception."""
    ds = vb.synth.make_dataset("extra_dd_fbx_wb", lambda m, t: vb.Pulsar(m, t), ntoa=300)
    psr = vb.Pulsar(ds.model, ds.toas)
    W = ds.walkers(4, seed=1)
    W[2, ds.info["free_names"].index("ECC")] = 1.5
    out = psr.lnlike_vectorized(W)
    assert out[2] == -np.inf and np.all(np.isfinite(out[[0, 1, 3]]))
    assert oracle.lnlike_batch(vb.ModelDescriptor(ds.model, ds.toas), W)[2] == -np.inf


def test_chi2_rejected_for_woodbury(vb, gp):
    g, psr = gp
    with pytest.raises(vb.VelaB200Error, match="white-noise"):
        psr.chi2(g.x0)


def test_api_function_mirrors(vb, golden):
    """The reference's own call patterns (test_NGC6440E.jl:126-214) through the mirrored names."""
    g = golden("NGC6440E")
    model, toas = g.model, g.toas
    calc_lnlike_s = vb.get_lnlike_serial_func(model, toas)
    calc_lnlike_p = vb.get_lnlike_parallel_func(model, toas)
    parv = model.read_param_values_to_vector()
    assert calc_lnlike_s(parv) == calc_lnlike_p(parv)
    assert vb.calc_chi2_reduced(model, toas, parv) < 1.2
    res = vb.form_residuals(model, toas, parv)
    assert np.all(np.abs(res) < 3 * toas.error)
    y, ninv = vb._calc_resids_and_Ninvdiag(model, toas, parv)
    assert len(y) == len(ninv) and np.all(ninv > 0)
    calc_lnpost_vec = vb.get_lnpost_func(model, toas, True)
    paramss = np.stack([parv, parv, parv])
    out = calc_lnpost_vec(paramss)
    assert out[0] == out[1] == out[2] and np.isfinite(out[0])
    assert np.isclose(out[0], vb.get_lnpost_func(model, toas)(parv), rtol=1e-15)
    psr = vb.Pulsar(model, toas)
    assert vb.calc_lnlike(psr, parv) == vb.calc_lnlike_serial(psr, parv) == vb.calc_lnlike_vectorized(psr, paramss)[0]
    assert vb.calc_lnpost(psr, parv) == vb.calc_lnpost_vectorized(psr, paramss)[0]
    cube = np.full(len(parv), 0.5)
    assert np.isfinite(vb.calc_lnprior(psr, vb.prior_transform(psr, cube)))


@pytest.mark.parametrize("name,ntoa", [("config3_gp", 800), ("extra_ecorr_gp", None), ("config2_ell1", 600)])
def test_batches_larger_than_the_scratch_run_in_passes(name, ntoa, vb, monkeypatch):
    """Host batches beyond the per-pass scratch budget are processed in passes (forced here with a 128-walker chunk)."""
    ds = vb.synth.make_dataset(name, lambda m, t: vb.Pulsar(m, t), ntoa=ntoa)
    W = ds.walkers(700, seed=12)
    whole = vb.Pulsar(ds.model, ds.toas).lnlike_vectorized(W)
    monkeypatch.setenv("VELA_B200_MAX_CHUNK", "128")
    chunked = vb.Pulsar(ds.model, ds.toas)
    monkeypatch.delenv("VELA_B200_MAX_CHUNK")
    got = chunked.lnlike_vectorized(W)
    assert np.all(np.isfinite(got))
    np.testing.assert_allclose(got, whole, rtol=1e-12)


@pytest.mark.parametrize("chunk", ["4", "8", "64"])
def test_ecorr_correction_staging_sizes(chunk, vb, monkeypatch):
    """The ECORR basis kernel stages TOAs in passes: several groups per pass, one group per pass, and (4 TOAs per pass
    against 8-TOA epochs) groups cut into pieces must all give the same correction."""
    ds = vb.synth.make_dataset("extra_ecorr_gp", lambda m, t: vb.Pulsar(m, t))
    psr = vb.Pulsar(ds.model, ds.toas)
    W = ds.walkers(70, seed=5)
    ref = psr.lnlike_vectorized(W)
    monkeypatch.setenv("VELA_B200_ECORR_CHUNK", chunk)
    got = psr.lnlike_vectorized(W)
    monkeypatch.delenv("VELA_B200_ECORR_CHUNK")
    assert np.all(np.isfinite(ref))
    np.testing.assert_allclose(got, ref, rtol=1e-11)


def test_shared_reciprocal_division_is_ieee(vb):
    """div3_by (three quotients from one reciprocal refinement, the proper-motion normalisation of solarsystem.jl:45-64),
    div_fast and rcp_fast (the division sequence without its special-case scaffolding, operands 2^-130 .. 2^130) must give
    the IEEE-rounded result bit for bit: 5 x 2^26 quotients, none may differ."""
    assert vb.lib.lib().vela_b200_debug_div3(1 << 26, 12345) == 0
