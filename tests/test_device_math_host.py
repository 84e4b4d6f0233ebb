"""The chain kernel's arithmetic, compiled for the CPU, against the oracle (`-m "not gpu"`).

tests/host_emul builds vela_device.cuh / vela_chain.cuh (the product's device source) with g++ behind a shim of the
CUDA builtins and runs, thread by thread, what the kernels run per walker and per (TOA, walker).  The oracle follows
the reference's operation order in full Double64; the kernel source takes documented shortcuts (small phase terms in
FP64, error-free leading-word subtraction, shared reciprocal) - this test bounds what they cost, without a GPU:
residuals within 1e-15 s (the parity budget is 1e-9 s), white-noise ln L within 1e-12 relative (budget 1e-9) for the
strict build; 1e-12 s and 2e-10 for the fast build (fused multiply-adds, series for the unit-vector normalisation).
"""
import numpy as np
import pytest

from conftest import GOLDEN_NAMES
from helpers import emul_chain, oracle_eval_factory

SYNTH = ["config2_ell1", "config3_gp", "config4_dd_wb", "extra_ell1_fbx", "extra_dd_fbx_wb", "extra_dmjump_bits_wb",
         "extra_ell1h_fd_wavex", "extra_ell1k_swx", "extra_ddk_wb", "extra_ddh", "extra_dds", "extra_ecorr",
         "extra_free_gp", "extra_free_gp_wb", "extra_mtm_only", "config5_array"]


# (residual [s], ln L relative) bounds of the two builds of the chain.  strict mirrors the reference operation for
# operation.  fast (the product's default, kFlagFast in vela_device.cuh) keeps the 500 s Roemer delay, every term added to
# it and the Double64 phase rounded exactly as strict does; it fuses multiply-adds in the small terms and evaluates
# chromatic powers / FD logarithms from a tabulated ln(frequency), which changes those (<= millisecond) terms by a few
# 1e-15 of themselves - enough to flip the last bit of the delay accumulator (5.7e-14 s) on a TOA in a thousand for
# models that carry such components, and nothing else.  (The bounds below are set by extra_ell1_fbx, whose synthetic CMX
# values make chromatic delays of hundreds of seconds: there one ulp of pow() - any two libm's differ by that - is 1e-13 s.)
BOUNDS = {False: (1e-15, 1e-12), True: (5e-13, 3e-10)}


def _check(vb, oracle, md, walkers, white):
    for fast in (False, True):
        tol_r, tol_l = BOUNDS[fast]
        y, w, chi, lg = emul_chain(md, walkers, emit=1, fast=fast)
        for b in range(len(walkers)):
            yo, wo = oracle.residuals(md, walkers[b])
            nt = len(md.toas)
            assert np.max(np.abs(y[b, :nt] - yo[:nt])) < tol_r, (fast, np.max(np.abs(y[b, :nt] - yo[:nt])))
            np.testing.assert_allclose(w[b], wo, rtol=4e-16)
            if md.n_rows > nt:   # wideband DM residuals [Hz-scaled DM units]
                np.testing.assert_allclose(y[b, nt:], yo[nt:], rtol=1e-12, atol=1e-12 * np.max(np.abs(yo[nt:])))
        if white:
            y0, w0, chi0, lg0 = emul_chain(md, walkers, emit=0, fast=fast)
            want = oracle.lnlike_batch(md, walkers)
            np.testing.assert_allclose(-(chi0 + lg0) / 2, want, rtol=tol_l)


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_reference_fixtures(name, vb, oracle, golden):
    g = golden(name)
    white = g.md.desc.kernel_kind == 0
    _check(vb, oracle, g.md, g.walkers[:6], white)


@pytest.mark.parametrize("name", SYNTH)
def test_synthetic_cases(name, vb, oracle):
    cfg = vb.synth.CONFIGS[name]
    ds = vb.synth.make_dataset(name, oracle_eval_factory(vb, oracle), ntoa=min(cfg["ntoa"], 600))
    md = vb.ModelDescriptor(ds.model, ds.toas)
    _check(vb, oracle, md, ds.walkers(4, seed=3), md.desc.kernel_kind == 0)


def test_phase_mode_matches_oracle(vb, oracle, golden):
    g = golden("NGC6440E")
    hi, lo, f, _, _ = emul_chain(g.md, g.x0[None, :], emit=2)
    ohi, olo, of, _, _ = oracle.phases(g.md, g.x0)
    nt = len(g.md.toas)
    # Double64 phase residual: hi + lo agree to 1e-13 cycle (phase_small joins in FP64)
    assert np.max(np.abs((hi[0, :nt] - ohi) + (lo[0, :nt] - olo))) < 1e-13
    np.testing.assert_allclose(f[0, :nt], of, rtol=1e-15)


@pytest.mark.parametrize("name", ["config3_gp", "config4_dd_wb", "config5_array", "extra_ell1h_fd_wavex", "extra_free_gp_wb"])
def test_speculative_evaluation_is_the_branching_one(name, vb, oracle):
    """The per-model fast build evaluates every gated shortcut unconditionally and defers the gates to a `redo` flag
    (Corr::spec); wherever the flag stays down the outputs must be the branching evaluation's, bit for bit, and walkers
    near the truth must practically never raise it.  A walker far from the default sky position must raise it."""
    from helpers import emul_spec_check
    cfg = vb.synth.CONFIGS[name]
    ds = vb.synth.make_dataset(name, oracle_eval_factory(vb, oracle), ntoa=min(cfg["ntoa"], 400))
    md = vb.ModelDescriptor(ds.model, ds.toas)
    for emit in (0, 1, 3):
        pairs, redo, bad = emul_spec_check(md, ds.walkers(6, seed=11), emit=emit)
        assert bad == 0 and redo <= 0.01 * pairs, (emit, pairs, redo, bad)
    sky = [i for i, n in enumerate(ds.info["free_names"]) if n in ("RAJ", "DECJ", "ELONG", "ELAT")]
    assert sky, ds.info["free_names"]
    far = ds.walkers(1, seed=2).copy()
    far[0, sky[0]] += 1e-3          # 200 arcsec off: the Shapiro expansion point is out of reach
    pairs, redo, bad = emul_spec_check(md, far, emit=1)
    assert bad == 0 and redo > 0


@pytest.mark.parametrize("name", ["config3_gp", "config2_ell1", "extra_ell1k_swx"])
def test_spin_phase_difference_route_and_its_gate(name, vb, oracle, monkeypatch):
    """kFlagDelta (fast build): the spin phase as a difference from its default-parameter value agrees with the oracle's
    Double64 phase to the same 1e-15 s as the double-double route, and a walker whose spin frequency is 7e-4 Hz off - 1e5
    cycles over the span, beyond the route's gate on most TOAs - falls back to the double-double code: still the oracle's
    residuals (hundreds of seconds) to 1e-12 s."""
    from helpers import emul_spec_check
    ds = vb.synth.make_dataset(name, oracle_eval_factory(vb, oracle), ntoa=500)
    md = vb.ModelDescriptor(ds.model, ds.toas)
    names = list(ds.info["free_names"])
    spin = [i for i, n in enumerate(names) if n in ("F0", "F1")]
    W = ds.walkers(3, seed=6)
    W[1, spin[0]] += 7e-4
    nt = len(md.toas)
    y, w, chi, lg = emul_chain(md, W, emit=1, fast=True)
    monkeypatch.setenv("VELA_B200_DELTA", "0")
    ydd, _, _, _ = emul_chain(md, W, emit=1, fast=True)
    monkeypatch.delenv("VELA_B200_DELTA")
    assert np.any(y[0, :nt] != ydd[0, :nt])                      # the route is in use ...
    for b in range(3):
        yo, _ = oracle.residuals(md, W[b])
        tol = 1e-12 if b == 1 else 1e-15
        assert np.max(np.abs(y[b, :nt] - yo[:nt])) < tol, (b, np.max(np.abs(y[b, :nt] - yo[:nt])))
    pairs, redo, bad = emul_spec_check(md, W, emit=1)
    assert bad == 0 and 0.2 * nt < redo <= nt                    # ... and its gate trips for the walker that is far off
