"""Pins the CPU oracle against the reference's own datasets and test bounds (`-m "not gpu"`).

The fixtures in tests/golden/ are the reference's test/datafiles/*.jlso re-encoded by
tests/golden/make_golden.py.  The reference tests hold bounds and identities, not golden numbers
(SURVEY §4); those bounds are restated here with their file:line, plus PINT's own CHI2 recorded in the
par file, plus the numbers recorded when the fixtures were generated (regression pin).
"""
import numpy as np
import pytest

from conftest import GOLDEN_NAMES


def test_ngc6440e_residual_and_chi2_bounds(golden, oracle):
    g = golden("NGC6440E")
    y, w = oracle.residuals(g.md, g.x0)
    assert len(g.toas) == 62                                              # test_NGC6440E.jl:35
    assert np.all(np.abs(y) < 3 * g.toas.error)                           # test_NGC6440E.jl:120-124
    assert np.all(w > 0)                                                  # test_NGC6440E.jl:130
    chi2 = oracle.chi2(g.md, g.x0)
    dof = len(g.toas) - len(g.x0)
    assert chi2 / dof < 1.2                                               # test_NGC6440E.jl:143
    # PINT's post-fit statistics for the same par/tim: CHI2 62.000000016, CHI2R 1.148 (NGC6440E.par:19-20)
    assert abs(chi2 - 62.000000016) < 5e-3
    assert abs(chi2 / dof - 1.1481) < 1e-3


def test_ngc6440e_model_structure(golden):
    g = golden("NGC6440E")
    assert set(g.model.get_free_param_names()) == {"F0", "F1", "PHOFF", "RAJ", "DECJ", "DM", "EFAC1", "EQUAD1"}  # :61-62
    assert g.model.get_num_timing_params() == 6                           # :63
    assert [type(c).__name__ for c in g.model.components] == [
        "SolarSystem", "DispersionTaylor", "Spindown", "PhaseOffset", "MeasurementNoise"]   # :88-108
    assert not g.model.components[0].ecliptic_coordinates and not g.model.components[0].planet_shapiro
    assert g.model.pulsar_name == "J1748-2021E" and g.model.ephem == "DE440"                # :7-10
    assert g.model.tzr_toa.index == 0 and g.model.tzr_toa.pulse_number == (0.0, 0.0)         # :48-52


def test_ngc6440e_lnlike_drops_when_noise_param_doubled(golden, oracle):
    g = golden("NGC6440E")
    p1 = g.x0.copy()
    p1[-3] *= 2                                                           # test_NGC6440E.jl:163-165
    assert oracle.lnlike(g.md, p1) < oracle.lnlike(g.md, g.x0)


def test_sim_sw_wb_bounds(golden, oracle):
    g = golden("sim_sw_wb")
    n = len(g.toas)
    assert n == 500 and g.toas.wideband
    y, w = oracle.residuals(g.md, g.x0)
    assert np.all(np.abs(y[:n]) < 4.1 * g.toas.error)                     # test_sim_sw_wb.jl:94
    assert np.all(np.abs(y[n:]) < 3.5 * g.toas.dm_error)                  # test_sim_sw_wb.jl:95
    # the reference's bounds are tight to the true maxima: they pin the wideband chain, not just its sign
    assert np.max(np.abs(y[:n]) / g.toas.error) > 3.9
    assert np.max(np.abs(y[n:]) / g.toas.dm_error) > 3.4
    assert oracle.chi2(g.md, g.x0) / (2 * n - len(g.x0)) < 1.5            # test_sim_sw_wb.jl:108
    assert [type(c).__name__ for c in g.model.components] == [
        "SolarSystem", "SolarWindDispersion", "DispersionTaylor", "ExclusiveDispersionJump", "Spindown",
        "PhaseOffset", "DispersionMeasurementNoise"]                      # test_sim_sw_wb.jl:70-90
    p1 = g.x0.copy()
    p1[-1] *= 2                                                           # test_sim_sw_wb.jl:121-123
    assert oracle.lnlike(g.md, p1) < oracle.lnlike(g.md, g.x0)


@pytest.mark.parametrize("name,ntoa,ngp,rank", [("sim3_gp", 2000, 3, 180), ("sim_dmgp_wb", 500, 2, 248)])
def test_woodbury_fixtures(name, ntoa, ngp, rank, golden, oracle):
    g = golden(name)
    assert len(g.toas) == ntoa                                            # test_sim3_gp.jl:28, test_sim_dmgp_wb.jl:28
    assert len(g.model.kernel.gp_components) == ngp                       # test_sim3_gp.jl:114, test_sim_dmgp_wb.jl:94
    assert g.md.n_basis == rank
    y, w = oracle.residuals(g.md, g.x0)
    assert len(y) == len(w) == g.md.n_rows and np.all(w > 0)              # test_sim3_gp.jl:117-122
    l0 = oracle.lnlike(g.md, g.x0)
    assert np.isfinite(l0)
    p1 = g.x0.copy()
    p1[-3] *= 2                                                           # test_sim3_gp.jl:132-134, test_sim_dmgp_wb.jl:112-114
    assert oracle.lnlike(g.md, p1) < l0


def test_sim3_gp_free_parameter_names(golden):
    g = golden("sim3_gp")
    assert set(g.model.get_free_param_names()) == {
        "RAJ", "DECJ", "PHOFF", "DM", "DM1", "DM2", "CM", "CM1", "CM2", "F0", "F1", "TNDMAMP", "TNDMGAM",
        "TNREDAMP", "TNREDGAM", "TNCHROMAMP", "TNCHROMGAM"}               # test_sim3_gp.jl:54-72
    assert g.model.get_num_timing_params() == 11                          # test_sim3_gp.jl:73


def test_woodbury_equals_dense_gaussian(golden, oracle):
    """ln L = -1/2 (y^T C^-1 y + log det C), C = N + M Phi M^T (test_woodbury_kernel.jl:87-88) on a real dataset."""
    g = golden("sim3_gp")
    y, w = oracle.residuals(g.md, g.x0)
    phiinv = oracle.noise_weights_inv(g.md, g.x0)
    Mb = np.asarray(g.model.kernel.noise_basis)
    # scale to O(1) before the dense solve: C is badly scaled in seconds^2
    s = np.sqrt(w)
    C = np.eye(len(y)) + (Mb * s[:, None]) @ np.diag(1 / phiinv) @ (Mb * s[:, None]).T
    ys = y * s
    sign, logdet = np.linalg.slogdet(C)
    dense = -0.5 * (ys @ np.linalg.solve(C, ys) + logdet - np.sum(np.log(w)))
    assert sign > 0
    np.testing.assert_allclose(oracle.lnlike(g.md, g.x0), dense, rtol=1e-8)


def test_sim2_ecorr_bounds(golden, oracle):
    """EcorrKernel dataset: test/test_sim2.jl:87-136 and PINT's CHI2 1999.9595 (pyvela/examples/sim2.par:17-18)."""
    g = golden("sim2")
    assert len(g.toas) == 2000 and type(g.model.kernel).__name__ == "EcorrKernel"
    assert set(g.model.get_free_param_names()) == {"RAJ", "DECJ", "PHOFF", "F0", "F1", "EFAC1", "ECORR1"}
    y, w = oracle.residuals(g.md, g.x0)
    assert np.all(np.abs(y) < 7 * g.toas.error)                           # test_sim2.jl:106-110
    chi2 = oracle.chi2(g.md, g.x0)
    assert chi2 / (len(g.toas) - len(g.x0)) < 1.2                         # test_sim2.jl:119
    assert abs(chi2 - 1999.9595338710205) < 0.02                          # PINT's GLS chi^2 incl. ECORR
    p1 = g.x0.copy()
    p1[-1] *= 2                                                           # test_sim2.jl:133-135
    assert oracle.lnlike(g.md, p1) < oracle.lnlike(g.md, g.x0)


def test_ecorr_likelihood_equals_dense_gaussian(golden, oracle):
    """White + ECORR ln L = -1/2 (r^T C^-1 r + log det C), C = N + U diag(ECORR^2) U^T (Johnson+ 2024), on sim2."""
    g = golden("sim2")
    y, w = oracle.residuals(g.md, g.x0)
    ecorr = g.model.read_params(g.x0)[g.model.param_handler.offset("ECORR")]
    total = 0.0
    for start, stop, idx in g.model.kernel.ecorr_groups:
        sl = slice(int(start) - 1, int(stop))
        n = sl.stop - sl.start
        C = np.diag(1 / w[sl]) + (ecorr**2 if idx else 0.0) * np.ones((n, n))
        total += y[sl] @ np.linalg.solve(C, y[sl]) + np.linalg.slogdet(C)[1]
    np.testing.assert_allclose(oracle.lnlike(g.md, g.x0), -0.5 * total, rtol=1e-10)


def test_gls_ecorr_identity(oracle):
    """test/test_woodbury_kernel.jl:96-160: WoodburyKernel{EcorrKernel} == dense Gaussian with C = Nc + M Phi M^T."""
    rng = np.random.default_rng(3)
    y, Nd, M, Ph = rng.standard_normal(9), 1 + rng.random(9), rng.standard_normal((9, 5)), 1 + rng.random(5)
    groups, ec = [[1, 3, 1], [4, 6, 2], [7, 9, 3]], np.array([0.5, 1.1, 0.9])
    U = np.kron(np.eye(3), np.ones((1, 3)))
    C = np.diag(Nd) + U.T @ np.diag(ec**2) @ U + M @ np.diag(1 / Ph) @ M.T
    want = -0.5 * (y @ np.linalg.solve(C, y) + np.linalg.slogdet(C)[1])
    assert oracle.gls_lnlike_raw(M, 1 / Nd, Ph, y, groups, ec) == pytest.approx(want, rel=1e-12)


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_oracle_matches_recorded_numbers(name, golden, oracle):
    """Regression pin: the numbers stored in the fixture by make_golden.py."""
    g = golden(name)
    y, w = oracle.residuals(g.md, g.x0)
    np.testing.assert_allclose(y, g.y, rtol=0, atol=1e-13)
    np.testing.assert_allclose(w, g.w, rtol=1e-13)
    np.testing.assert_allclose(oracle.lnlike_batch(g.md, g.walkers), g.lnlike, rtol=1e-10)
    np.testing.assert_allclose(oracle.lnlike(g.md, g.x0), g.extra["lnlike_default"], rtol=1e-12)
    # serial == threaded (test_NGC6440E.jl:159): 1 thread vs all threads
    np.testing.assert_array_equal(oracle.lnlike_batch(g.md, g.walkers[:8], nthreads=1),
                                  oracle.lnlike_batch(g.md, g.walkers[:8], nthreads=4))
