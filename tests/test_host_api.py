"""Host-side logic and the C-ABI surface (`-m "not gpu"`: no compute calls, no GPU needed)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, has_gpu


def test_library_loads_and_exports_every_declared_symbol(vb):
    hdr = open(os.path.join(ROOT, "include", "vela_b200.h")).read()
    declared = sorted(set(re.findall(r"VELA_API[^;(]*?\b(vela_b200_\w+)\s*\(", hdr)))
    assert len(declared) >= 20
    lib = ctypes.CDLL(vb.lib.LIB_PATH)
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/vela_b200.h but not exported"
    assert sorted(vb.lib.EXPORTS) == declared


def test_no_cpu_fallback_fails_loudly(vb, golden):
    if has_gpu():
        pytest.skip("a GPU is present")
    g = golden("NGC6440E")
    with pytest.raises(vb.VelaB200Error, match="no CPU fallback"):
        vb.Pulsar(g.model, g.toas)
    h = ctypes.c_void_p()
    rc = vb.lib.lib().vela_b200_create(g.md.byref(), ctypes.byref(h))
    assert rc == -2 and b"no CPU path" in vb.lib.lib().vela_b200_last_error()      # VELA_ERR_NO_DEVICE


def test_create_rejects_bad_descriptors(vb, golden):
    g = golden("NGC6440E")
    L = vb.lib.lib()
    h = ctypes.c_void_p()
    md = vb.ModelDescriptor(g.model, g.toas)
    md.desc.abi_version = 99
    assert L.vela_b200_create(md.byref(), ctypes.byref(h)) == -1
    assert b"ABI version" in L.vela_b200_last_error()
    md = vb.ModelDescriptor(g.model, g.toas)
    md.desc.toa_stride_bytes = 100
    assert L.vela_b200_create(md.byref(), ctypes.byref(h)) == -1
    md = vb.ModelDescriptor(g.model, g.toas)
    md._tzr[30:31].view(np.uint64)[0] = 7        # not a TZR TOA (timing_model.jl:37)
    assert L.vela_b200_create(md.byref(), ctypes.byref(h)) == -1
    assert L.vela_b200_create(None, ctypes.byref(h)) == -1


def test_param_handler_ordering(vb):
    """parameter.jl:87-119: non-noise singles, non-noise multis, noise singles, noise multis."""
    M = vb.model
    P = M.Parameter
    ph = M.ParamHandler(
        [P("A", 1.0, 0, True), P("NZ", 9.0, 0, False, noise=True), P("B", 2.0, 0, False)],
        [M.MultiParameter("E", [P("E1", 7.0, 0, False, noise=True), P("E2", 8.0, 0, True, noise=True)]),
         M.MultiParameter("F", [P("F0", 3.0, 0, False), P("F1", 4.0, 0, True)])])
    np.testing.assert_array_equal(ph._default_values, [1.0, 2.0, 3.0, 4.0, 9.0, 7.0, 8.0])
    np.testing.assert_array_equal(ph._free_indices, [2, 3, 5, 6])            # 1-based, as in Julia
    assert ph.get_free_param_names() == ["B", "F0", "NZ", "E1"]
    assert ph.get_free_param_prefixes() == ["B", "F", "NZ", "E"]
    assert ph.get_num_timing_params() == 2
    full = ph.read_params([20.0, 30.0, 90.0, 70.0])
    np.testing.assert_array_equal(full, [1.0, 20.0, 30.0, 4.0, 90.0, 70.0, 8.0])
    np.testing.assert_array_equal(ph.read_param_values_to_vector(full), [20.0, 30.0, 90.0, 70.0])
    with pytest.raises(AssertionError):
        ph.read_params([1.0])                                                 # parameter.jl:162


def test_fixture_param_handler_matches_reference(golden):
    """_default_values of NGC6440E as decoded from the reference file."""
    g = golden("NGC6440E")
    ph = g.model.param_handler
    assert len(ph._default_values) == 16 and ph._nfree == 8
    np.testing.assert_array_equal(ph._free_indices, [3, 4, 7, 10, 12, 13, 15, 16])
    assert ph.offset("F_") == 8 and ph.offset("DM") == 9 and ph.offset("F") == 11
    assert ph.offset("EFAC") == 14 and ph.offset("EQUAD") == 15


def test_descriptor_offsets(vb, golden):
    g = golden("sim_sw_wb")
    d = g.md.desc
    assert d.n_components == 7 and d.wideband == 1 and d.toa_stride_bytes == 264 and d.n_toas == 500
    kinds = [d.components[i].kind for i in range(d.n_components)]
    D = vb.descriptor
    assert kinds == [D.COMP_SOLAR_SYSTEM, D.COMP_SOLAR_WIND, D.COMP_DISPERSION_TAYLOR, D.COMP_DM_JUMP_EXCLUSIVE,
                     D.COMP_SPINDOWN, D.COMP_PHASE_OFFSET, D.COMP_DM_MEASUREMENT_NOISE]
    ss = d.components[0]
    assert ss.flags == (D.FLAG_ECLIPTIC | D.FLAG_PLANET_SHAPIRO)
    ph = g.model.param_handler
    assert [ss.par[i] for i in range(6)] == [ph.offset(n) for n in ("ELONG", "ELAT", "PMELONG", "PMELAT", "PX", "POSEPOCH")]
    sp = d.components[4]
    assert sp.par[0] == ph.offset("F_") and sp.par[1] == ph.offset("F") and sp.len[0] == 3
    assert d.n_index_masks == 3
    g3 = golden("sim3_gp").md.desc
    assert g3.kernel_kind == D.KERNEL_WOODBURY_WHITE and g3.n_gp == 3 and g3.basis_cols == 180 and g3.basis_rows == 2000
    assert [g3.gp[i].kind for i in range(3)] == [D.GP_POWERLAW_RED, D.GP_POWERLAW_DM, D.GP_POWERLAW_CHROM]


def test_ecorr_descriptor(vb, golden):
    g = golden("sim2")
    d = g.md.desc
    assert d.kernel_kind == vb.descriptor.KERNEL_ECORR and d.n_ecorr_groups == 250
    assert [d.ecorr_groups[i] for i in range(6)] == [1, 8, 1, 9, 16, 1]
    assert d.par_ecorr == g.model.param_handler.offset("ECORR")
    bad = g.model.kernel.ecorr_groups.copy()
    bad[3, 0] += 1                     # hole in the TOA partition
    M = vb.model
    m2 = M.TimingModel(g.model.pulsar_name, g.model.ephem, g.model.clock, g.model.units, g.model.epoch,
                       g.model.components, M.EcorrKernel(bad), g.model.param_handler, g.model.tzr_toa, g.model.priors)
    with pytest.raises(ValueError, match="partition"):
        vb.ModelDescriptor(m2, g.toas)


def test_priors_host_side(vb, golden):
    g = golden("NGC6440E")
    pri = vb.priors
    assert np.isfinite(pri.lnprior(g.model.priors, g.x0))                       # test_NGC6440E.jl:172
    x = g.x0.copy()
    x[0] += 1.0                                                                # outside the RAJ box
    assert pri.lnprior(g.model.priors, x) == -np.inf
    batch = np.stack([g.x0, x])
    np.testing.assert_array_equal(np.isfinite(pri.lnprior(g.model.priors, batch)), [True, False])
    cube = pri.prior_transform(g.model.priors, np.full(len(g.x0), 0.5))
    assert np.all(np.isfinite(cube)) and np.isfinite(pri.lnprior(g.model.priors, cube))   # :177-179,198
    # quantile/logpdf consistency of the default families and the custom ones (known_priors.jl)
    from scipy import integrate
    for d in (pri.Uniform(-1, 3), pri.Normal(0.5, 2.0), pri.LogNormal(0.0, 0.25), pri.LogUniform(1e-3, 10.0),
              pri.KINPriorDistribution(), pri.SINIPriorDistribution(),
              pri.LatitudePriorDistribution(), pri.PXPriorDistribution(2.0)):
        lo, hi = float(d.quantile(0.1)), float(d.quantile(0.7))
        mass, _ = integrate.quad(lambda t: float(np.exp(d.logpdf(t))), lo, hi)
        assert mass == pytest.approx(0.6, abs=1e-6), type(d).__name__
    # STIGMA: the mirror keeps the reference's quantile q/(2-q) verbatim (known_priors.jl:61) although it is not the
    # inverse of its cdf 2s^2/(1+s^2) (:60) — drop-in behaviour over self-consistency
    assert float(pri.STIGMAPriorDistribution().quantile(0.5)) == pytest.approx(0.5 / 1.5)
    assert float(np.exp(pri.STIGMAPriorDistribution().logpdf(0.5))) == pytest.approx(4 * 0.5 / 1.25**2)


def test_npz_round_trip(vb, golden, tmp_path):
    g = golden("sim_dmgp_wb")
    path = str(tmp_path / "x.npz")
    vb.readwrite.save_pulsar_npz(path, g.model, g.toas)
    m2, t2, _ = vb.readwrite.load_pulsar_npz(path)
    np.testing.assert_array_equal(t2.records, g.toas.records)
    np.testing.assert_array_equal(m2.param_handler._default_values, g.model.param_handler._default_values)
    np.testing.assert_array_equal(np.asarray(m2.kernel.noise_basis), np.asarray(g.model.kernel.noise_basis))
    assert [type(c).__name__ for c in m2.components] == [type(c).__name__ for c in g.model.components]
    assert m2.get_marginalized_param_names()[:2] == ["PLREDSIN_0001", "PLREDSIN_0002"]   # gp_noise.jl:87-97


def test_shard_range(vb):
    sr = vb.parallel.shard_range
    for n in (0, 1, 7, 8192, 8193):
        for world in (1, 2, 3, 8):
            blocks = [sr(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))


def test_chain_specialisation_source_and_nvrtc_compile(vb, oracle):
    """The per-model translation unit names every component once and compiles with NVRTC without a GPU."""
    import ctypes as C
    from helpers import oracle_eval_factory
    L = vb.lib.lib()
    ds = vb.synth.make_dataset("extra_ell1_fbx", oracle_eval_factory(vb, oracle), ntoa=120)
    md = vb.ModelDescriptor(ds.model, ds.toas)
    buf = C.create_string_buffer(1 << 16)
    n = L.vela_b200_debug_spec_source(md.byref(), buf, len(buf))
    src = buf.value.decode()
    assert n == len(src) and src.count("VELA_APPLY(") == len(ds.model.components)
    assert "vela_chain_spec" in src and "kSpecProg" in src
    size = L.vela_b200_debug_spec_compile(md.byref())
    if size < 0 and b"libnvrtc not available" in L.vela_b200_last_error():
        pytest.skip("libnvrtc not installed")
    assert size > 10000, L.vela_b200_last_error()


@pytest.mark.parametrize("name,ntoa", [("config3_gp", 600), ("config4_dd_wb", 400), ("config5_array", 500),
                                       ("extra_ecorr_gp", None), ("fixture:sim3_gp", None), ("fixture:sim_dmgp_wb", None)])
def test_structured_sigma_analysis_matches_dense(name, ntoa, vb, oracle, golden):
    """Host half of the structured Sigma path: the harmonic structure is detected on the numbers of the basis
    (synthetic bases and the PINT-made bases of the reference fixtures) and the column-sum table reproduces
    M^T N^-1 M to rounding."""
    import ctypes as C
    from helpers import oracle_eval_factory
    if name.startswith("fixture:"):
        g = golden(name.split(":")[1])
        model, toas, x = g.model, g.toas, g.x0
    else:
        ds = vb.synth.make_dataset(name, oracle_eval_factory(vb, oracle), ntoa=ntoa)
        model, toas, x = ds.model, ds.toas, ds.truth
    md = vb.ModelDescriptor(model, toas)
    M = np.asarray(model.kernel.noise_basis)
    p = M.shape[1]
    _, w = oracle.residuals(md, x)
    out = np.zeros(p * (p + 1) // 2)
    info = (C.c_int32 * 3)()
    dp = C.POINTER(C.c_double)
    n_r = vb.lib.lib().vela_b200_debug_structured_sigma(md.byref(), w.ctypes.data_as(dp), out.ctypes.data_as(dp), info)
    assert 0 < n_r <= p * (p + 1) // 4 and info[2] >= 1          # at least 2x fewer rows than the packed triangle
    dense = M.T @ (M * w[:, None])
    il = np.tril_indices(p)
    scale = np.sqrt(np.outer(np.diag(dense), np.diag(dense)))[il]
    assert np.max(np.abs(out - dense[il]) / scale) < 1e-13


def test_structured_sigma_rejects_unstructured_basis(vb, oracle):
    """A basis that is not harmonic (random columns) stays on the dense path."""
    from helpers import oracle_eval_factory
    ds = vb.synth.make_dataset("config3_gp", oracle_eval_factory(vb, oracle), ntoa=300)
    k = ds.model.kernel
    rng = np.random.default_rng(0)
    bad = vb.model.WoodburyKernel(k.inner_kernel, k.gp_components, rng.standard_normal(k.noise_basis.shape))
    model = vb.model.TimingModel(ds.model.pulsar_name, ds.model.ephem, ds.model.clock, ds.model.units, ds.model.epoch,
                                 ds.model.components, bad, ds.model.param_handler, ds.model.tzr_toa, ds.model.priors)
    md = vb.ModelDescriptor(model, ds.toas)
    assert vb.lib.lib().vela_b200_debug_structured_sigma(md.byref(), None, None, None) == 0


def test_structured_sigma_on_perturbed_bases(vb, oracle):
    """The structure is verified on the numbers of M, never assumed: per-row scales (zero and negative ones
    included), a group with its own phase, a noisy column and log-spaced harmonics must either be served exactly
    through the table or leave the handle on the dense path."""
    import ctypes as C
    from helpers import oracle_eval_factory
    M_ = vb.model
    vb.synth.CONFIGS["_tmp_small_gp"] = dict(ntoa=400, span_d=3000.0, red=6, dmgp=5, chromgp=4, nback=2, walkers=16, seed=31)
    try:
        ds = vb.synth.make_dataset("_tmp_small_gp", oracle_eval_factory(vb, oracle))
    finally:
        del vb.synth.CONFIGS["_tmp_small_gp"]
    k = ds.model.kernel
    n, p = k.noise_basis.shape
    rng = np.random.default_rng(3)
    x = rng.uniform(-np.pi, np.pi, n)
    w = rng.uniform(0.5, 2.0, n) * 1e12
    dp = C.POINTER(C.c_double)
    L = vb.lib.lib()

    def basis(scales, phase_shift=(0.0, 0.0, 0.0)):
        cols = []
        for (nh, s, ph) in zip((6, 5, 4), scales, phase_shift):
            kk = np.arange(1, nh + 1)
            cols += [s[:, None] * np.sin(np.outer(x + ph, kk)), s[:, None] * np.cos(np.outer(x + ph, kk))]
        return np.concatenate(cols, axis=1)

    def run(Mat, gps=None):
        kern = M_.WoodburyKernel(k.inner_kernel, gps or k.gp_components, Mat)
        model = M_.TimingModel(ds.model.pulsar_name, ds.model.ephem, ds.model.clock, ds.model.units, ds.model.epoch,
                               ds.model.components, kern, ds.model.param_handler, ds.model.tzr_toa, ds.model.priors)
        md = vb.ModelDescriptor(model, ds.toas)
        out = np.zeros(p * (p + 1) // 2)
        info = (C.c_int32 * 3)()
        n_r = L.vela_b200_debug_structured_sigma(md.byref(), w.ctypes.data_as(dp), out.ctypes.data_as(dp), info)
        if n_r > 0:
            dense = Mat.T @ (Mat * w[:, None])
            il = np.tril_indices(p)
            scale = np.sqrt(np.outer(np.diag(dense), np.diag(dense)))[il] + 1e-300
            assert np.max(np.abs(out - dense[il]) / scale) < 1e-12
        return n_r, info[2]

    s1 = np.ones(n)
    s2 = rng.uniform(0.5, 8.0, n)
    s2[::7] = 0.0                                   # rows where the second group vanishes
    s3 = -s2 * s2                                   # a negative scale
    n_r, groups = run(basis((s1, s2, s3)))
    assert n_r > 0 and groups == 3
    # second group on its own phase: it cannot share the table; whatever path is chosen must stay exact
    n_r2, groups2 = run(basis((s1, s2, s3), phase_shift=(0.0, 0.3, 0.0)))
    assert groups2 <= 2 and (n_r2 == 0 or n_r2 > n_r)
    # one noisy column (1e-9 relative): its group is demoted
    noisy = basis((s1, s2, s3))
    noisy[:, 3] *= 1 + 1e-9 * rng.standard_normal(n)
    n_r3, groups3 = run(noisy)
    assert groups3 <= 2
    # log-spaced harmonics (non-integer multiples of the fundamental) are generic columns
    gps = (M_.PowerlawRedNoiseGP(4, 2, 2.0), k.gp_components[1], k.gp_components[2])
    js = np.exp(np.asarray(gps[0].ln_js, dtype=np.float64))
    red = np.concatenate([np.sin(np.outer(x, js)), np.cos(np.outer(x, js))], axis=1)
    rest = basis((s1, s2, s3))[:, 12:]
    n_r4, groups4 = run(np.concatenate([red, rest], axis=1), gps=gps)
    assert n_r4 == 0 or groups4 >= 2


def test_priors_reference_literals(vb):
    """The literal cases of test/test_priors.jl:3-44 (SimplePrior on PHOFF, truncated Normal on F0, Uniform on F1)."""
    P = vb.priors
    priors = [P.SimplePrior("PHOFF", P.Uniform(-0.5, 0.5), P.USER_DEFINED_PRIOR),
              P.SimplePriorMulti("F", 1, P.TruncatedNormal(100.0, 1e-7, 0.0, np.inf), P.USER_DEFINED_PRIOR),
              P.SimplePriorMulti("F", 2, P.Uniform(-1.01e-14, -0.9e-14), P.USER_DEFINED_PRIOR)]
    good, bad = np.array([0.1, 100.0, -1e-14]), np.array([1.1, -200.0, -2e-14])
    assert np.isfinite(P.lnprior(priors, good)) and not np.isfinite(P.lnprior(priors, bad))
    for k in range(3):
        assert np.isfinite(priors[k].distribution.logpdf(good[k])) and not np.isfinite(priors[k].distribution.logpdf(bad[k]))
    assert float(priors[0].distribution.quantile(0.5)) == 0.0
    assert float(priors[1].distribution.quantile(0.5)) == pytest.approx(100.0)
    assert float(priors[2].distribution.quantile(0.0)) == -1.01e-14
    np.testing.assert_allclose(P.prior_transform(priors, [0.5, 0.5, 0.0]), [0.0, 100.0, -1.01e-14], rtol=1e-12, atol=1e-30)
    d = vb.descriptor.prior_descriptors(priors)
    assert d is not None and [d[k].kind for k in range(3)] == [vb.descriptor.PRIOR_UNIFORM, vb.descriptor.PRIOR_TRUNCATED_NORMAL,
                                                             vb.descriptor.PRIOR_UNIFORM]


@pytest.mark.parametrize("name,ntoa", [("config3_gp", 6000), ("config4_dd_wb", 14000), ("config5_array", 6000), ("config3_gp_ecorr", 6000)])
def test_two_stage_column_sums_match_direct_sums(name, ntoa, vb, oracle):
    """The two-stage evaluation of the structured path's column sums (cell moments with 32 Chebyshev columns per scale
    function, then constant stage-2 matrices: Jacobi-Anger expansion of exp(i m x) around each cell's centre) against
    the direct sums over the table T, on the host (vela_b200_debug_two_stage): agreement to ~1e-14 of the largest sum."""
    import ctypes as C
    from helpers import oracle_eval_factory
    ds = vb.synth.make_dataset(name, oracle_eval_factory(vb, oracle), ntoa=ntoa)
    md = vb.ModelDescriptor(ds.model, ds.toas)
    L = vb.lib.lib()
    dp = C.POINTER(C.c_double)
    info = (C.c_int64 * 8)()
    nc = int(L.vela_b200_debug_two_stage(md.byref(), None, None, None, None, info))
    assert nc > 0 and info[1] == 1, list(info)
    y, w = oracle.residuals(md, ds.truth)
    wy = w * y
    Rd, Rt = np.zeros(nc), np.zeros(nc)
    L.vela_b200_debug_two_stage(md.byref(), w.ctypes.data_as(dp), wy.ctypes.data_as(dp), Rd.ctypes.data_as(dp),
                                Rt.ctypes.data_as(dp), info)
    n_u = (md.n_basis + 31) // 32 * 32
    nr = nc - n_u
    assert np.max(np.abs(Rd[:nr] - Rt[:nr])) <= 5e-14 * np.max(np.abs(Rd[:nr]))
    # u = M^T (w y) sums terms of both signs: bound every column by its sum of magnitudes, not by the cancelled result
    Mb = np.asarray(ds.model.kernel.noise_basis)
    bound = np.abs(Mb).T @ np.abs(wy)
    assert np.all(np.abs(Rd[nr:nr + md.n_basis] - Rt[nr:nr + md.n_basis]) <= 2e-14 * bound)
    assert info[4] * 32 < 0.75 * nc      # fewer stage-1 columns than table columns: that is the point
