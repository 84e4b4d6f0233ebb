"""The NVRTC-specialised chain kernel against the generic interpreting kernel: same source, same flags, so every
output must be bit-identical; and both stay within the oracle tolerances (covered by the other GPU tests, which
run with the specialised kernel by default)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [("config2_ell1", 600), ("config3_gp", 700), ("config4_dd_wb", 400), ("extra_ell1_fbx", None),
         ("extra_dd_fbx_wb", None), ("extra_ell1h_fd_wavex", None), ("extra_ddk_wb", None), ("extra_ecorr_gp", None),
         ("extra_free_gp_wb", None)]


@pytest.mark.parametrize("name,ntoa", CASES)
def test_specialized_chain_is_bit_identical(name, ntoa, vb):
    ds = vb.synth.make_dataset(name, lambda m, t: vb.Pulsar(m, t), ntoa=ntoa)
    psr = vb.Pulsar(ds.model, ds.toas)
    assert psr.chain_specialized(), f"no specialised kernel: {psr.jit_log()}"
    W = ds.walkers(70, seed=5)
    y1, w1 = psr.resids_and_Ninvdiag(ds.truth)
    p1 = psr.phases(ds.truth)
    l1 = psr.lnlike_vectorized(W)
    assert not psr.chain_specialized(False)
    y0, w0 = psr.resids_and_Ninvdiag(ds.truth)
    p0 = psr.phases(ds.truth)
    l0 = psr.lnlike_vectorized(W)
    assert psr.chain_specialized(True)
    np.testing.assert_array_equal(y1, y0)
    np.testing.assert_array_equal(w1, w0)
    for a, b in zip(p1, p0):
        np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(l1, l0)
    assert np.all(np.isfinite(l1))


def test_fixture_with_specialized_kernel(golden, vb):
    g = golden("sim_sw_wb")
    psr = vb.Pulsar(g.model, g.toas)
    assert psr.chain_specialized()
    a = psr.lnlike_vectorized(g.walkers)
    psr.chain_specialized(False)
    np.testing.assert_array_equal(a, psr.lnlike_vectorized(g.walkers))


@pytest.mark.parametrize("name,ntoa", [("config3_gp", 700), ("config4_dd_wb", 400), ("extra_free_gp_wb", None)])
def test_speculative_chain_redoes_walkers_outside_the_gates(name, ntoa, vb):
    """The per-model fast build evaluates straight-line and re-evaluates, per warp, the walkers for which a gate failed
    (chain_body, Corr::spec).  A batch that mixes posterior-width walkers with walkers far off the default sky position
    (the Shapiro expansion and the unit-vector shortcut are out of reach there) must come out bit-identical to the
    generic kernel, which only knows the branching code - flagged and unflagged walkers side by side in every group."""
    ds = vb.synth.make_dataset(name, lambda m, t: vb.Pulsar(m, t), ntoa=ntoa)
    psr = vb.Pulsar(ds.model, ds.toas)
    assert psr.chain_specialized(), f"no specialised kernel: {psr.jit_log()}"
    W = ds.walkers(96, seed=8)
    sky = [i for i, n in enumerate(ds.info["free_names"]) if n in ("RAJ", "DECJ", "ELONG", "ELAT")]
    assert sky
    rng = np.random.default_rng(4)
    far = rng.choice(len(W), 29, replace=False)
    W[far, sky[0]] += rng.uniform(3e-4, 3e-2, len(far)) * rng.choice([-1.0, 1.0], len(far))
    l1 = psr.lnlike_vectorized(W)
    y1, w1 = psr.resids_and_Ninvdiag(W[far[0]])
    assert not psr.chain_specialized(False)
    l0 = psr.lnlike_vectorized(W)
    y0, w0 = psr.resids_and_Ninvdiag(W[far[0]])
    assert psr.chain_specialized(True)
    np.testing.assert_array_equal(l1, l0)
    np.testing.assert_array_equal(y1, y0)
    np.testing.assert_array_equal(w1, w0)
