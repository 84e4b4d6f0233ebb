"""An INDEPENDENT restatement of the NGC6440E chain in numpy.longdouble (80-bit), written directly from
the reference formulas, used to pin the double-double oracle "two independent ways" (SURVEY §7 step 1).

Chain (test/test_NGC6440E.jl:88-108): SolarSystem(equatorial, no planets) -> DispersionTaylor -> Spindown ->
PhaseOffset -> MeasurementNoise.  longdouble has a 64-bit mantissa: phases of ~3e9 cycles keep ~3e-10 cycle.
"""
import numpy as np
import pytest

ld = np.longdouble
pytestmark = pytest.mark.skipif(np.finfo(ld).nmant < 63, reason="needs an 80-bit long double")


def _chain_longdouble(model, rec, P, names, tzr):
    """Returns (phase, spin_frequency, doppler, delay) for one TOA record, all in longdouble."""
    g = lambda n: ld(P[names[n][0]])                                   # noqa: E731
    t = ld(rec[0]) + ld(rec[1])
    R, V, S = rec[6:9].astype(ld), rec[9:12].astype(ld), rec[12:15].astype(ld)
    # solarsystem.jl:45-64,67-134 (PM = 0 => Lhat = x0; PX = 0)
    a, d = g("RAJ"), g("DECJ")
    assert P[names["PMRA"][0]] == 0 and P[names["PMDEC"][0]] == 0 and P[names["PX"][0]] == 0
    L = np.array([np.cos(a) * np.cos(d), np.sin(a) * np.cos(d), np.sin(d)], dtype=ld)
    AU, MSUN = ld(499.00478383615643), ld(4.92549094830932e-06)
    delay = -(L @ R) - 2 * MSUN * np.log((np.sqrt(S @ S) - L @ S) / AU)
    doppler = L @ V
    # dispersion.jl:15-26, component.jl:70-79
    o, n = names["DM"]
    tt = t - delay
    dm = ld(P[o]) + (ld(P[o + 1]) * (tt - g("DMEPOCH")) if n > 1 else 0)
    nu = ld(rec[3]) * (1 - doppler)
    delay = delay + dm / (nu * nu)
    # spindown.jl:15-33
    dt = t - delay
    o, n = names["F"]
    F = [ld(P[o + i]) for i in range(n)]
    phase = g("F_") * dt
    fact = ld(1)
    f = g("F_")
    for i, Fi in enumerate(F):
        fact *= (i + 1)
        phase += Fi * dt ** (i + 1) / fact
        f += Fi * dt ** i / (fact / (i + 1))
    # phase_offset.jl:12-13
    if not tzr:
        phase -= g("PHOFF")
    return phase, f, doppler


def test_ngc6440e_chi2_longdouble_vs_oracle(golden, oracle):
    g = golden("NGC6440E")
    ph = g.model.param_handler
    names = ph._offsets
    P = ph.read_params(g.x0)
    tzr_phase, _, _ = _chain_longdouble(g.model, g.model.tzr_toa.record(), P, names, tzr=True)
    efac = ld(P[names["EFAC"][0]])
    equad = ld(P[names["EQUAD"][0]])
    chi2 = ld(0)
    lnl = ld(0)
    res = []
    for rec in g.toas.records:
        phase, f, doppler = _chain_longdouble(g.model, rec, P, names, tzr=False)
        dphase = phase - (ld(rec[4]) + ld(rec[5])) - tzr_phase
        r = dphase / (f * (1 + doppler))
        s2 = (ld(rec[2]) ** 2 + equad**2) * efac**2
        chi2 += r * r / s2
        lnl += r * r / s2 + np.log(s2)
        res.append(float(r))
    y, _ = oracle.residuals(g.md, g.x0)
    # residuals: longdouble keeps ~1e-10 cycle / 61 Hz ~ 2e-12 s
    assert np.max(np.abs(np.array(res) - y)) < 2e-11
    assert float(chi2) == pytest.approx(oracle.chi2(g.md, g.x0), rel=1e-8)
    assert float(-lnl / 2) == pytest.approx(oracle.lnlike(g.md, g.x0), rel=1e-9)
