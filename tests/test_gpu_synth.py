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
         ("extra_ddh", None, 48), ("extra_dds", None, 48), ("extra_ecorr", None, 64), ("extra_ecorr_gp", None, 64)]


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
                                       ("extra_ecorr_gp", None)])
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
