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


@pytest.mark.parametrize("name,ntoa", [("config3_gp", 900), ("config2_ell1", 700), ("extra_ell1_fbx", None)])
def test_spin_phase_difference_route(name, ntoa, vb, oracle, monkeypatch):
    """kFlagDelta (fast build): the spin phase as a difference from its default-parameter value, against the double-double
    route (VELA_B200_DELTA=0) and the oracle - also for walkers whose spin frequency is so far off that |phase - phase0|
    exceeds the route's 1e4-cycle gate and the double-double code takes over (in line / through the redo loop)."""
    ds = vb.synth.make_dataset(name, lambda m, t: vb.Pulsar(m, t), ntoa=ntoa)
    W = ds.walkers(40, seed=6)
    names = list(ds.info["free_names"])
    spin = [i for i, n in enumerate(names) if n in ("F0", "F1")]      # the first element of the F vector, whatever it is called
    assert spin, names
    far = [3, 17, 31]
    W[far, spin[0]] += 7e-4                                             # ~1e5 cycles over the data span
    psr = vb.Pulsar(ds.model, ds.toas)
    l1 = psr.lnlike_vectorized(W)
    y1, w1 = psr.resids_and_Ninvdiag(W[0])
    yf, _ = psr.resids_and_Ninvdiag(W[far[0]])
    monkeypatch.setenv("VELA_B200_DELTA", "0")
    ref = vb.Pulsar(ds.model, ds.toas)
    monkeypatch.delenv("VELA_B200_DELTA")
    l0 = ref.lnlike_vectorized(W)
    y0, w0 = ref.resids_and_Ninvdiag(W[0])
    yf0, _ = ref.resids_and_Ninvdiag(W[far[0]])
    nt = len(ds.toas)
    assert np.max(np.abs(y1[:nt] - y0[:nt])) < 1e-15
    np.testing.assert_array_equal(w1, w0)
    # beyond the gate (most TOAs of these walkers) both handles run the double-double code; residuals of hundreds of seconds
    assert np.max(np.abs(yf[:nt])) > 1.0 and np.max(np.abs(yf[:nt] - yf0[:nt])) < 1e-12
    np.testing.assert_allclose(l1, l0, rtol=1e-11)
    yfo, _ = oracle.residuals(vb.ModelDescriptor(ds.model, ds.toas), W[far[0]])
    assert np.max(np.abs(yf[:nt] - yfo[:nt])) < 1e-9
    md = vb.ModelDescriptor(ds.model, ds.toas)
    yo, _ = oracle.residuals(md, W[0])
    assert np.max(np.abs(y1[:nt] - yo[:nt])) < 1e-12
    np.testing.assert_allclose(l1[:6], oracle.lnlike_batch(md, W[:6]), rtol=1e-9)
