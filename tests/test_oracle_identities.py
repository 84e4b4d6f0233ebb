"""Known-answer identities the reference's unit tests assert, applied to the oracle (`-m "not gpu"`)."""
import numpy as np
import pytest


def test_mikkola_round_trip(oracle):
    """u == mikkola(u - e sin u, e) for the 18 (u, e) pairs of test/test_dd.jl:2-11."""
    us = [-np.pi / 4, 0.0, np.pi / 4, 3 * np.pi / 4, 4 * np.pi / 3, 7 * np.pi / 3]
    for e in (0.0, 0.5, 0.3):
        for u in us:
            assert oracle.mikkola(u - e * np.sin(u), e) == pytest.approx(u, rel=1e-8, abs=1e-12)


def test_sincos_correctly_rounded(oracle):
    """The oracle's sky-position sin/cos equals the exactly rounded value (mpmath, 200 bits)."""
    import mpmath
    mpmath.mp.prec = 200
    rng = np.random.default_rng(11)
    for x in np.concatenate([rng.uniform(-7, 7, 400), rng.uniform(-500, 500, 50), [0.0, 4.663868852922854, -0.3553169573849282]]):
        s, c = oracle.sincos_cr(float(x))
        assert s == float(mpmath.sin(mpmath.mpf(float(x)))) and c == float(mpmath.cos(mpmath.mpf(float(x))))


def test_gls_lnlike_identities(oracle):
    """test/test_woodbury_kernel.jl:37-92 with the same shapes (100 x 5, random)."""
    rng = np.random.default_rng(5)
    y = rng.standard_normal(100)
    Ndiag = 1 + rng.random(100)
    M = rng.standard_normal((100, 5))
    Phiinv = 1 + rng.random(5)
    got = oracle.gls_lnlike_raw(M, 1 / Ndiag, Phiinv, y)
    Sigmainv = np.diag(Phiinv) + M.T @ (M / Ndiag[:, None])
    u = M.T @ (y / Ndiag)
    brute = -0.5 * (y @ (y / Ndiag) - u @ np.linalg.solve(Sigmainv, u) + np.sum(np.log(Ndiag))
                    - np.sum(np.log(Phiinv)) + np.linalg.slogdet(Sigmainv)[1])
    assert got == pytest.approx(brute, rel=1e-12)
    C = np.diag(Ndiag) + M @ np.diag(1 / Phiinv) @ M.T
    assert got == pytest.approx(-0.5 * (y @ np.linalg.solve(C, y) + np.linalg.slogdet(C)[1]), rel=1e-12)


def test_gls_not_positive_definite_is_minus_inf(oracle):
    """isposdef(Sigma^-1) == false  =>  -Inf  (gls_likelihood.jl:211-213)."""
    M = np.ones((4, 2))
    assert oracle.gls_lnlike_raw(M, np.ones(4), np.array([-10.0, -10.0]), np.ones(4)) == -np.inf


def test_powerlaw_weights_closed_form(vb, oracle):
    """test/test_gp_noise.jl:23-34: weights_inv[j] == 1 / powerlaw(10^A, gamma, exp(ln_j) f1, f1); 1e-6 for the
    red-noise GP because its ln_js are Float32 (gp_noise.jl:111), exact-ish for DM/chromatic (:59-68,94-103)."""
    M = vb.model
    f1, amp, gam = 1e-8, -13.5, 3.5
    for cls, tol in ((M.PowerlawRedNoiseGP, 1e-6), (M.PowerlawDispersionNoiseGP, 1e-12), (M.PowerlawChromaticNoiseGP, 1e-12)):
        gp = cls(3, 2, 2.0)
        assert len(gp.ln_js) == 5
        P1 = oracle.powerlaw(10.0**amp, gam, f1, f1)
        w1 = 1 / (P1 * np.exp(-gam * gp.ln_js))
        exact_ln_js = np.concatenate([(1 / 2.0) ** np.arange(2, 0, -1), np.log(np.arange(1, 4))])
        w2 = np.array([1 / oracle.powerlaw(10.0**amp, gam, np.exp(l) * f1, f1) for l in exact_ln_js])
        assert np.all(np.abs(w1 / w2 - 1) < tol)
    assert M.PowerlawRedNoiseGP(3, 2, 2.0).ln_js[3] == float(np.float32(np.log(2.0)))
    assert M.PowerlawDispersionNoiseGP(3, 2, 2.0).ln_js[3] == np.log(2.0)


def test_noise_weights_inv_blocks(golden, oracle):
    """Phi^-1 = [sin block, cos block] per GP component, blocks equal (test_woodbury_kernel.jl:27-35)."""
    g = golden("sim3_gp")
    phi = oracle.noise_weights_inv(g.md, g.x0)
    assert len(phi) == 180 and np.all(phi > 0)
    for k in range(3):
        np.testing.assert_array_equal(phi[60 * k: 60 * k + 30], phi[60 * k + 30: 60 * k + 60])


def _single_toa_model(vb, components, single, multi, toa_value=0.0, index=1, freq=2.5e9):
    """The literal TOA of test/runtests.jl:13-59 with a caller-chosen component tuple."""
    M = vb.model
    eph = M.SolarSystemEphemeris((18.0354099, 450.01472245, 195.05827732), (-9.96231954e-05, 3.31555854e-06, 1.12968547e-06),
                                 (-15.89483533, -450.17185232, -195.18212616), (-1610.1796849, -1706.87348483, -681.22381513),
                                 (-2392.85651431, 3109.13626083, 1405.71912274), (140.85922773, -217.65571843, -74.64804201),
                                 (9936.62957939, -3089.07377113, -1486.17339104), (11518.60924426, -9405.0693235, -4126.91030657))
    toa = M.TOA((toa_value, 0.0), 1e-6, freq, (0.0, 0.0), eph, index)
    tzr = M.make_tzr_toa((5 * 86400.0, 0.0), 2.5e9, eph)
    ph = M.ParamHandler([M.Parameter(n, v, 0, True) for n, v in single],
                        [M.MultiParameter(n, [M.Parameter(f"{n}{i}", v, 0, True) for i, v in enumerate(vs)]) for n, vs in multi])
    model = M.TimingModel("T", "DE440", "c", "TDB", 0.0, tuple(components), M.WhiteNoiseKernel(), ph, tzr, ())
    return vb.ModelDescriptor(model, M.TOAs.from_list([toa]))


def test_component_identities(vb, oracle):
    M = vb.model
    spin = [("F_", 100.0)], [("F", [0.0, -1e-14])]
    # Spindown alone at t = 0: zero phase, f = 100 Hz (test_spindown.jl:14-18)
    md = _single_toa_model(vb, [M.Spindown()], *spin)
    hi, lo, f, delay, tzr = oracle.phases(md, np.zeros(0))
    assert f[0] == 100.0 and delay[0] == 0.0
    # DM delay = DM / nu^2 (test_dispersion.jl:17-19): barycentric TOA phase shifts by -F * DM/nu^2
    md2 = _single_toa_model(vb, [M.DispersionTaylor(), M.Spindown()], [("DMEPOCH", 0.0), ("F_", 100.0)],
                            [("DM", [1e16, 0.0]), ("F", [0.0, 0.0])])
    _, _, _, delay2, _ = oracle.phases(md2, np.zeros(0))
    assert delay2[0] == pytest.approx(1e16 / 2.5e9**2, rel=1e-15)
    # ChromaticTaylor with index 2 equals the dispersion delay for CM = DM * 1e-12 (test_chromatic.jl:17-20)
    md3 = _single_toa_model(vb, [M.ChromaticTaylor(), M.Spindown()], [("CMEPOCH", 0.0), ("TNCHROMIDX", 2.0), ("F_", 100.0)],
                            [("CM", [1e16 * 1e-12, 0.0]), ("F", [0.0, 0.0])])
    _, _, _, delay3, _ = oracle.phases(md3, np.zeros(0))
    assert delay3[0] == pytest.approx(delay2[0], rel=1e-12)
    # PHOFF: applies to TOAs, not to the TZR TOA (phase_offset.jl:12-13): phase residual moves by -PHOFF
    a = _single_toa_model(vb, [M.Spindown(), M.PhaseOffset()], [("PHOFF", 0.0), ("F_", 100.0)], [("F", [0.0, 0.0])])
    b = _single_toa_model(vb, [M.Spindown(), M.PhaseOffset()], [("PHOFF", 0.25), ("F_", 100.0)], [("F", [0.0, 0.0])])
    pa, pb = oracle.phases(a, np.zeros(0)), oracle.phases(b, np.zeros(0))
    assert pb[0][0] - pa[0][0] == pytest.approx(-0.25, abs=1e-9) and pa[4][0] == pb[4][0]
    # JUMP phase = JUMP * F (test_jump.jl:19): compare with the same model whose mask selects no TOA
    c1 = _single_toa_model(vb, [M.Spindown(), M.ExclusivePhaseJump(np.array([1], dtype=np.uint32))],
                           [("F_", 100.0)], [("JUMP", [2e-3]), ("F", [0.5, 0.0])])
    c0 = _single_toa_model(vb, [M.Spindown(), M.ExclusivePhaseJump(np.array([0], dtype=np.uint32))],
                           [("F_", 100.0)], [("JUMP", [2e-3]), ("F", [0.5, 0.0])])
    assert oracle.phases(c1, np.zeros(0))[0][0] - oracle.phases(c0, np.zeros(0))[0][0] == pytest.approx(2e-3 * 100.5, abs=1e-7)
    # SolarSystem: delay changes, barycentred TOA untouched (test_solarsystem.jl:41-51)
    ss_par = [("POSEPOCH", 0.0), ("PX", 3e-12), ("RAJ", 1.2), ("DECJ", 1.25), ("PMRA", -7e-16), ("PMDEC", -5e-16), ("F_", 100.0)]
    d = _single_toa_model(vb, [M.SolarSystem(False, True), M.Spindown()], ss_par, [("F", [0.0, 0.0])])
    _, _, fd, delay_d, _ = oracle.phases(d, np.zeros(0))
    assert abs(delay_d[0]) > 1.0 and fd[0] != 100.0
    # glitch before its epoch contributes nothing (test_glitch.jl:32)
    gl = [("GLEP_", [1e6]), ("GLPH_", [0.39]), ("GLF0_", [1.7e-6]), ("GLF1_", [-6e-15]), ("GLF2_", [0.0]),
          ("GLF0D_", [0.0]), ("GLTD_", [0.0]), ("F", [0.0, 0.0])]
    e = _single_toa_model(vb, [M.Spindown(), M.Glitch()], [("F_", 100.0)], gl)
    assert oracle.phases(e, np.zeros(0))[0][0] == pytest.approx(pa[0][0] + 0.0, abs=1e-3)


def test_free_amplitude_powerlaw_gp_delay(vb, oracle):
    """Un-marginalised power-law GPs as delay components with the parameters of test/test_gp_noise.jl:5-12,41-48,
    75-83 (the reference asserts only `isfinite(delay)`); here the delay is checked against a direct mpmath
    evaluation of gp_noise.jl:23-58 with exact trigonometry instead of the complex recurrence."""
    import mpmath
    mpmath.mp.prec = 120
    M = vb.model
    amp, gam, f1, epoch, t = -13.5, 3.5, 1e-8, -1.2e8, 3.3e7
    sins = {"PLRED": (1.3, 2.0, 1.1, 0.12, -0.23), "PLDM": (3.0, 2.0, 1.1, 0.12, -0.23), "PLCHROM": (1.3, 2.0, 1.1, 0.12, -0.23)}
    coss = {"PLRED": (-1.6, 2.3, 1.5, -1.11, -0.8), "PLDM": (3.0, 2.0, 1.5, -1.11, -0.8), "PLCHROM": (1.3, -2.0, 1.5, -1.11, -0.8)}
    nu = 2.5e9
    for cls, tn, pre in ((M.PowerlawRedNoiseGP, "TNRED", "PLRED"), (M.PowerlawDispersionNoiseGP, "TNDM", "PLDM"),
                         (M.PowerlawChromaticNoiseGP, "TNCHROM", "PLCHROM")):
        gp = cls(3, 2, 2.0)
        single = [(tn + "AMP", amp), (tn + "GAM", gam), (pre + "FREQ", f1), (pre + "EPOCH", epoch), ("F_", 100.0)]
        if pre == "PLCHROM":
            single.append(("TNCHROMIDX", 4.0))
        md = _single_toa_model(vb, [gp, M.Spindown()], single,
                               [(pre + "SIN_", sins[pre]), (pre + "COS_", coss[pre]), ("F", [0.0, 0.0])], toa_value=t, freq=nu)
        assert md.desc.components[0].len[1] == 2          # two log-spaced harmonics precede ln(1) == 0
        delay = oracle.phases(md, np.zeros(0))[3][0]
        fyr = mpmath.mpf(1) / 3600 / 24 / mpmath.mpf("365.25")
        P1 = mpmath.mpf(10) ** (2 * amp) / (12 * mpmath.pi**2 * fyr**3) * (fyr / f1) ** gam * f1
        phi1 = 2 * mpmath.pi * f1 * (mpmath.mpf(t) - epoch)
        tot = mpmath.mpf(0)
        for i, (lnj, j) in enumerate(zip(gp.ln_js, gp.js())):
            jj = mpmath.mpf(float(j)) if i < 2 else mpmath.mpf(i - 1)
            tot += mpmath.exp(-mpmath.mpf(gam) / 2 * float(lnj)) * (sins[pre][i] * mpmath.sin(jj * phi1) + coss[pre][i] * mpmath.cos(jj * phi1))
        want = mpmath.sqrt(P1) * tot
        if pre == "PLDM":
            want = want * mpmath.mpf(1.4e9) ** 2 / mpmath.mpf(nu) ** 2
        elif pre == "PLCHROM":
            want = want * mpmath.mpf(1400.0) ** 4 * (mpmath.mpf(1e6) / nu) ** 4
        assert np.isfinite(delay) and delay == pytest.approx(float(want), rel=1e-12)
    # the free-amplitude variance matches the marginalised prior weight: (sigma1 jfac)^2 == 1 / weights_inv
    red = M.PowerlawRedNoiseGP(3, 2, 2.0)
    P1 = oracle.powerlaw(10.0**amp, gam, f1, f1)
    np.testing.assert_allclose(P1 * np.exp(-(gam / 2) * red.ln_js) ** 2, P1 * np.exp(-gam * red.ln_js), rtol=1e-14)
