"""GPU parity on synthetic datasets of the BASELINE.json shapes and on component-coverage cases.

Tolerances (north star): residuals within 1 ns, ln-likelihood within 1e-9 relative, against the CPU oracle on
the same inputs.  Reduced sizes are used where the oracle must run every walker; the full BASELINE sizes are
covered by size-independent properties plus an oracle check on a walker subset (test_gpu_full_size.py).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [("config2_ell1", 1500, 96), ("config3_gp", 1200, 96), ("config4_dd_wb", 700, 64), ("config5_array", 900, 40),
         ("extra_ell1_fbx", None, 64), ("extra_dd_fbx_wb", None, 64), ("extra_dmjump_bits_wb", None, 64),
         ("extra_ell1h_fd_wavex", None, 48), ("extra_ell1k_swx", None, 48), ("extra_ddk_wb", None, 48),
         ("extra_ddh", None, 48), ("extra_dds", None, 48), ("extra_ecorr", None, 64), ("extra_ecorr_gp", None, 64),
         ("extra_free_gp", None, 48), ("extra_free_gp_wb", None, 48), ("extra_ecorr_gp_r80", None, 40),
         ("extra_ecorr_gp_r140", None, 40), ("extra_mtm_only", None, 48)]


@pytest.fixture(scope="module")
def datasets(vb):
    cache = {}

    def get(name, ntoa):
        if name not in cache:
            ds = vb.synth.make_dataset(name, lambda m, t: vb.Pulsar(m, t), ntoa=ntoa)
            cache[name] = (ds, vb.Pulsar(ds.model, ds.toas), vb.ModelDescriptor(ds.model, ds.toas))
        return cache[name]

    return get


@pytest.mark.parametrize("name,ntoa,nw", CASES)
def test_residuals_and_lnlike(name, ntoa, nw, datasets, oracle):
    ds, psr, md = datasets(name, ntoa)
    assert ds.info["connect_rms_s"] < 1e-12
    n = len(ds.toas)
    y, w = psr.resids_and_Ninvdiag(ds.truth)
    yo, wo = oracle.residuals(md, ds.truth)
    assert np.max(np.abs(y[:n] - yo[:n])) < 1e-9                      # 1 ns
    np.testing.assert_allclose(w, wo, rtol=1e-13)
    if ds.toas.wideband:
        assert np.max(np.abs(y[n:] - yo[n:]) * np.sqrt(wo[n:])) < 1e-9
    W = ds.walkers(nw, seed=21)
    got = psr.lnlike_vectorized(W)
    want = oracle.lnlike_batch(md, W)
    assert np.all(np.isfinite(want))
    np.testing.assert_allclose(got, want, rtol=1e-9)


@pytest.mark.parametrize("name,ntoa", [("config3_gp", 1200), ("config4_dd_wb", 700), ("config5_array", 900),
                                       ("extra_ecorr_gp", None), ("extra_ecorr_gp_r80", None), ("extra_ecorr_gp_r140", None)])
def test_sigma_and_u_match_oracle(name, ntoa, datasets, oracle):
    """_calc_Σinv__and__MT_Ninv_y parity (the quantity test_woodbury_kernel.jl:66-76 checks)."""
    ds, psr, md = datasets(name, ntoa)
    S, u = psr.sigma_and_u(ds.truth)
    So, uo = oracle.sigma(md, ds.truth)
    scale = np.sqrt(np.outer(np.diag(So), np.diag(So)))
    assert np.max(np.abs(S - So) / scale) < 1e-12
    np.testing.assert_allclose(u, uo, rtol=1e-9, atol=1e-9 * np.abs(uo).max())
    np.testing.assert_allclose(psr.noise_weights_inv(ds.truth), oracle.noise_weights_inv(md, ds.truth), rtol=1e-12)


def test_phases_api_matches_oracle(datasets, oracle):
    ds, psr, md = datasets("extra_ell1_fbx", None)
    hi, lo, f = psr.phases(ds.truth)
    ho, lo_o, fo, _, _ = oracle.phases(md, ds.truth)
    assert np.max(np.abs((hi - ho) + (lo - lo_o))) < 1e-9            # cycles
    np.testing.assert_allclose(f, fo, rtol=1e-14)


@pytest.mark.parametrize("name,ntoa", [("config3_gp", 1200), ("config4_dd_wb", 700), ("extra_ecorr_gp", None)])
def test_marginalized_mean_and_factor(name, ntoa, datasets, oracle):
    """ahat = Sigma u and the upper factor of Sigma^-1 (pyvela spnta.py:376-394) against scipy on the oracle's
    Sigma^-1, u; compared in the Sigma^-1 energy norm because Sigma^-1 is ill-conditioned by construction."""
    from scipy.linalg import cho_factor, cho_solve
    ds, psr, md = datasets(name, ntoa)
    W = ds.walkers(5, seed=3)
    ahat, cf, lnl = psr.marginalized_offset_mean_vectorized(W, with_cf=True, with_lnlike=True)
    # the conditional-mean outputs come from the CTA-per-walker factorisation, ln L of small ranks from the warp-per-walker
    # one: same Sigma^-1, another operation order
    np.testing.assert_allclose(lnl, psr.lnlike_vectorized(W), rtol=1e-13)
    for b in range(len(W)):
        So, uo = oracle.sigma(md, W[b])
        scale = np.sqrt(np.outer(np.diag(So), np.diag(So)))
        R = cf[b]
        assert np.all(np.tril(R, -1) == 0) and np.all(np.diag(R) > 0)
        assert np.max(np.abs(R.T @ R - So) / scale) < 1e-12
        want = cho_solve(cho_factor(So), uo)
        d = ahat[b] - want
        assert d @ So @ d < 1e-16 * (want @ So @ want)
    # single-sample mirrors of the SPNTA helpers
    a1, R1 = psr.get_marginalized_param_offset_mean_and_covinvcf(W[0])
    np.testing.assert_array_equal(a1, ahat[0])
    std = psr.get_marginalized_param_std(W[0])
    So, _ = oracle.sigma(md, W[0])
    np.testing.assert_allclose(std, np.sqrt(np.diag(np.linalg.inv(So))), rtol=1e-6)
    a, lnp = psr.get_marginalized_param_offset_sample(W[0], np.random.default_rng(0))
    z = R1 @ (a - a1)
    assert abs(lnp - (-0.5 * z @ z + np.log(np.diag(R1)).sum() - 0.5 * len(a) * np.log(2 * np.pi))) < 1e-8 * abs(lnp)
    assert psr.get_marginalized_gp_noise_realization(W[0], np.random.default_rng(0)).shape == (psr.n_rows,)


def test_marginalized_mean_refused_for_white_kernels(vb, datasets):
    ds, psr, md = datasets("config2_ell1", 1500)
    assert not psr.has_marginalized_gp_noise
    with pytest.raises(vb.VelaB200Error):
        psr.marginalized_offset_mean_vectorized(ds.walkers(2, seed=1))
    assert psr.get_marginalized_param_offset_mean(ds.truth).size == 0


def test_whitened_residuals(datasets, oracle):
    """spnta.py:523-556: the per-group closed form equals the reference's dense ECORR normal equations."""
    from scipy.linalg import cho_factor, cho_solve
    ds, psr, md = datasets("extra_ecorr", None)
    x = ds.walkers(1, seed=9)[0]
    y, w = psr.resids_and_Ninvdiag(x)
    groups = [g for g in np.asarray(ds.model.kernel.ecorr_groups, dtype=np.int64).reshape(-1, 3) if g[2] > 0]
    ph = ds.model.param_handler
    full = np.array(ph._default_values)
    full[np.asarray(ph._free_indices0)] = x
    U = np.zeros((len(y), len(groups)))
    phi = np.empty(len(groups))
    for k, (a0, a1, idx) in enumerate(groups):
        U[a0 - 1:a1, k] = 1.0
        phi[k] = full[ph.offset("ECORR") + idx - 1] ** 2
    NinvU = U * w[:, None]
    ahat = cho_solve(cho_factor(np.diag(1 / phi) + U.T @ NinvU), NinvU.T @ y)
    np.testing.assert_allclose(psr.whitened_time_residuals(x), y - U @ ahat, rtol=1e-9, atol=1e-15)
    np.testing.assert_allclose(psr.scaled_toa_uncertainties(x), 1 / np.sqrt(w), rtol=1e-15)
    # wideband + GP: the whitened residuals remove exactly M a for the draw made with the same generator
    ds4, psr4, _ = datasets("config4_dd_wb", 700)
    n = len(ds4.toas)
    real = psr4.get_marginalized_gp_noise_realization(ds4.truth, np.random.default_rng(5))
    np.testing.assert_array_equal(psr4.whitened_time_residuals(ds4.truth, np.random.default_rng(5)),
                                  psr4.time_residuals(ds4.truth) - real[:n])
    np.testing.assert_array_equal(psr4.whitened_dm_residuals(ds4.truth, np.random.default_rng(5)),
                                  psr4.dm_residuals(ds4.truth) - real[n:])


@pytest.mark.parametrize("name,ntoa", [("config3_gp", 1200), ("config4_dd_wb", 700), ("config5_array", 900),
                                       ("extra_ecorr_gp", None), ("extra_dmjump_bits_wb", None)])
def test_structured_and_dense_sigma_paths_agree(name, ntoa, datasets, vb, oracle, monkeypatch):
    """The column-sum (structured) contraction and the dense pair contraction are the same quantity: both run
    here on the same walkers, and both are checked against the oracle."""
    ds, psr, md = datasets(name, ntoa)
    assert psr.sigma_structured()
    monkeypatch.setenv("VELA_B200_STRUCTURED", "0")
    dense = vb.Pulsar(ds.model, ds.toas)
    monkeypatch.delenv("VELA_B200_STRUCTURED")
    assert not dense.sigma_structured()
    W = ds.walkers(150, seed=8)
    a, b = psr.lnlike_vectorized(W), dense.lnlike_vectorized(W)
    np.testing.assert_allclose(a, b, rtol=1e-11)
    np.testing.assert_allclose(b, oracle.lnlike_batch(md, W), rtol=1e-9)
    S1, u1 = psr.sigma_and_u(ds.truth)
    S0, u0 = dense.sigma_and_u(ds.truth)
    scale = np.sqrt(np.outer(np.diag(S0), np.diag(S0)))
    assert np.max(np.abs(S1 - S0) / scale) < 1e-13
    np.testing.assert_allclose(u1, u0, rtol=1e-10, atol=1e-10 * np.abs(u0).max())
    # small and odd batch sizes go through the split-K planner of both paths
    for B in (1, 3, 33):
        np.testing.assert_allclose(psr.lnlike_vectorized(W[:B]), a[:B], rtol=1e-12)
        np.testing.assert_allclose(dense.lnlike_vectorized(W[:B]), b[:B], rtol=1e-12)


def test_wideband_model_dm_and_scaled_errors(datasets, oracle):
    """spnta.py:605-645 mirrors on a wideband dataset: model DM = measured DM - DM residual, scaled DM errors."""
    ds, psr, md = datasets("config4_dd_wb", 700)
    n = len(ds.toas)
    yo, wo = oracle.residuals(md, ds.truth)
    np.testing.assert_allclose(psr.model_dm(ds.truth), (ds.toas.dm - yo[n:]) * 2.41e-16, rtol=1e-12)
    np.testing.assert_allclose(psr.scaled_dm_uncertainties(ds.truth), 1 / np.sqrt(wo[n:]), rtol=1e-13)
    with pytest.raises(NotImplementedError):
        datasets("config3_gp", 1200)[1].model_dm(datasets("config3_gp", 1200)[0].truth)


def test_warp_per_walker_cholesky_matches_cta_kernel(vb, monkeypatch):
    """chol_warp_kernel (opt-in: VELA_B200_CHOL_WARP=1; one warp per walker, layout-table assembly) against the default
    CTA-per-walker kernel on the same walkers, small and large batches (the small ones also take the wide slab reduction),
    and a rank that leaves the second row slot partly empty."""
    for name, ntoa in (("config3_gp", 1500), ("extra_dmjump_bits_wb", None)):
        ds = vb.synth.make_dataset(name, lambda m, t: vb.Pulsar(m, t), ntoa=ntoa)
        psr = vb.Pulsar(ds.model, ds.toas)
        if not psr.sigma_structured():
            continue
        for B in (3, 70, 1500):
            W = ds.walkers(B, seed=11)
            monkeypatch.delenv("VELA_B200_CHOL_WARP", raising=False)
            ref = psr.lnlike_vectorized(W)
            monkeypatch.setenv("VELA_B200_CHOL_WARP", "1")
            got = psr.lnlike_vectorized(W)
            monkeypatch.delenv("VELA_B200_CHOL_WARP", raising=False)
            assert np.all(np.isfinite(ref))
            np.testing.assert_allclose(got, ref, rtol=1e-13)


@pytest.mark.parametrize("name,ntoa", [("config3_gp", 1500), ("config4_dd_wb", 900), ("extra_mtm_only", None),
                                       ("extra_dmjump_bits_wb", None)])
def test_tensor_core_cholesky_matches_cta_kernel(name, ntoa, vb, oracle, monkeypatch):
    """chol_dmma_kernel (default for ranks <= 64 on the structured path: one warp per walker, 8 x 8 blocks in DMMA
    fragment registers) against the CTA-per-walker kernel (VELA_B200_CHOL_DMMA=0) on the same walkers, at batch sizes
    that take the graph replay, the deep split-K and the two-stage column sums, and against the oracle."""
    if name not in vb.synth.CONFIGS:
        pytest.skip("no such synthetic case")
    ds = vb.synth.make_dataset(name, lambda m, t: vb.Pulsar(m, t), ntoa=ntoa)
    psr = vb.Pulsar(ds.model, ds.toas)
    if not psr.sigma_structured():
        pytest.skip("dense Sigma path: the CTA kernel is the only one")
    for B in (1, 5, 70, 1500):
        W = ds.walkers(B, seed=13)
        got = psr.lnlike_vectorized(W)
        monkeypatch.setenv("VELA_B200_CHOL_DMMA", "0")
        ref = psr.lnlike_vectorized(W)
        monkeypatch.delenv("VELA_B200_CHOL_DMMA")
        assert np.all(np.isfinite(ref))
        np.testing.assert_allclose(got, ref, rtol=1e-12)
    md = vb.ModelDescriptor(ds.model, ds.toas)
    np.testing.assert_allclose(got[:4], oracle.lnlike_batch(md, W[:4]), rtol=1e-9)
    # a walker whose Sigma^-1 is not positive definite (NaN amplitude) must come out as -Inf from this kernel too
    bad = W[:3].copy()
    names = list(ds.info["free_names"])
    amp = [i for i, n in enumerate(names) if "AMP" in n.upper()]
    if amp:
        bad[1, amp[0]] = np.nan
        out = psr.lnlike_vectorized(bad)
        assert np.isfinite(out[0]) and out[1] == -np.inf and np.isfinite(out[2])


@pytest.mark.parametrize("name,ntoa", [("config3_gp", 1500), ("extra_ecorr_gp_r140", None), ("config4_dd_wb", 900)])
def test_panel_cholesky_matches_shared_memory_kernel(name, ntoa, vb, oracle, monkeypatch):
    """chol_panel_kernel (large ranks: 32-column panels + DMMA trailing updates on the global scratch) forced onto small
    ranks with VELA_B200_CHOL_SMEM=0, against the shared-memory kernel and the older global-scratch kernel."""
    ds = vb.synth.make_dataset(name, lambda m, t: vb.Pulsar(m, t), ntoa=ntoa)
    W = ds.walkers(67, seed=4)
    ref = vb.Pulsar(ds.model, ds.toas).lnlike_vectorized(W)
    monkeypatch.setenv("VELA_B200_CHOL_SMEM", "0")
    big = vb.Pulsar(ds.model, ds.toas)
    monkeypatch.delenv("VELA_B200_CHOL_SMEM")
    got = big.lnlike_vectorized(W)
    monkeypatch.setenv("VELA_B200_CHOL_PANEL", "0")
    old = big.lnlike_vectorized(W[:66])          # another batch size: not the cached graph of the panel kernel
    monkeypatch.delenv("VELA_B200_CHOL_PANEL")
    np.testing.assert_allclose(got, ref, rtol=1e-12)
    np.testing.assert_allclose(old, ref[:66], rtol=1e-12)
    md = vb.ModelDescriptor(ds.model, ds.toas)
    np.testing.assert_allclose(got[:4], oracle.lnlike_batch(md, W[:4]), rtol=1e-9)
    # a walker whose Sigma^-1 is not finite: -Inf for it alone (the panel kernel has no pivot test: NaN must carry through)
    gam = [i for i, n in enumerate(ds.info["free_names"]) if "GAM" in n.upper()]
    assert gam
    bad = W[:5].copy()
    bad[1, gam[0]] = np.nan
    bad[3, gam[0]] = 4000.0                      # Phi^-1 overflows
    out = big.lnlike_vectorized(bad)
    assert out[1] == -np.inf and not np.isnan(out).any()
    np.testing.assert_array_equal(out[[0, 2, 4]], got[[0, 2, 4]])


def test_panel_cholesky_small_rank(vb, oracle, monkeypatch):
    """The same kernel on a rank below its panel width (MarginalizedTimingModel alone, rank 4, dense Sigma path)."""
    ds = vb.synth.make_dataset("extra_mtm_only", lambda m, t: vb.Pulsar(m, t))
    W = ds.walkers(37, seed=2)
    ref = vb.Pulsar(ds.model, ds.toas).lnlike_vectorized(W)
    monkeypatch.setenv("VELA_B200_CHOL_SMEM", "0")
    big = vb.Pulsar(ds.model, ds.toas)
    monkeypatch.delenv("VELA_B200_CHOL_SMEM")
    np.testing.assert_allclose(big.lnlike_vectorized(W), ref, rtol=1e-12)
