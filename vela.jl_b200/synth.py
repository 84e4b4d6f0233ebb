"""Synthetic phase-connected pulsar datasets of the shapes named in BASELINE.json `configs`.

No PINT here: TOAs are built directly in Vela's internal representation (seconds from PEPOCH as a
double-double, light-second ephemeris vectors — pyvela/pyvela/toas.py:11-109 defines that contract).

Phase connection needs the model phase at each trial TOA in double-double precision.  The generator
does not compute it itself: it is handed an *evaluator factory* — `make_eval(model, toas)` returning
an object with `.phases(params) -> (hi, lo, f)` and `.resids_and_Ninvdiag(params) -> (y, w)`.
`bench.py` passes `vela_jl_b200.Pulsar` (the CUDA library evaluates its own inputs); the CPU tests
pass a wrapper around the oracle.  Either way the dataset is deterministic given (config, seed).
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional

import numpy as np

from . import model as M
from . import priors as P

DAY = 86400.0
AU = 499.00478383615643           # light-seconds (src/model/solarsystem.jl:5)
OBL = 0.4090926006005829
YEAR = 365.25 * DAY
DMCONST = 4.148808e15             # Hz^2 pc^-1 cm^3 s: DM [pc cm^-3] -> Vela units [Hz] (parameters.py:42-46)
F_YR = 1.0 / YEAR


@dataclass
class SynthDataset:
    name: str
    model: M.TimingModel
    toas: M.TOAs
    truth: np.ndarray                 # free-parameter vector the data were generated at
    sigma_par: np.ndarray             # per-parameter scatter used for walker batches
    noise_draw: Dict[int, tuple]      # free index -> ("lognormal"|"loguniform"|"normal", a, b)
    info: dict

    def walkers(self, n: int, seed: int = 0) -> np.ndarray:
        """Row b = truth + sigma_par * z_b (posterior-width scatter); noise
        hyper-parameters are drawn from narrow priors around the truth."""
        rng = np.random.default_rng(seed)
        out = self.truth[None, :] + self.sigma_par[None, :] * rng.standard_normal((n, len(self.truth)))
        for k, (kind, a, b) in self.noise_draw.items():
            if kind == "lognormal":
                out[:, k] = a * np.exp(b * rng.standard_normal(n))
            elif kind == "loguniform":
                out[:, k] = np.exp(rng.uniform(np.log(a), np.log(b), n))
            else:
                out[:, k] = a + b * rng.standard_normal(n)
        return np.ascontiguousarray(out)


# ---------------------------------------------------------------------------------------------
# toy solar-system ephemeris: circular, coplanar orbits in the ecliptic, rotated to equatorial axes
# ---------------------------------------------------------------------------------------------
_PLANETS = [("jupiter", 5.2044, 11.862), ("saturn", 9.5826, 29.4571), ("venus", 0.7233, 0.61519),
            ("uranus", 19.2184, 84.0205), ("neptune", 30.07, 164.8)]


def _orbit(t, radius_au, period_yr, phase0):
    ang = 2 * np.pi * t / (period_yr * YEAR) + phase0
    x, y = radius_au * AU * np.cos(ang), radius_au * AU * np.sin(ang)
    w = 2 * np.pi / (period_yr * YEAR)
    vx, vy = -radius_au * AU * w * np.sin(ang), radius_au * AU * w * np.cos(ang)
    ce, se = np.cos(OBL), np.sin(OBL)
    pos = np.stack([x, ce * y, se * y], axis=1)
    vel = np.stack([vx, ce * vy, se * vy], axis=1)
    return pos, vel


def ephemeris_words(t: np.ndarray, rng) -> np.ndarray:
    """24 ephemeris words per TOA (ssb_obs_pos, ssb_obs_vel, obs_sun_pos, 5 planets)."""
    earth, vel = _orbit(t, 1.0, 1.0, rng.uniform(0, 2 * np.pi))
    sun = 0.0 * earth
    sun[:, 0] += 1.5 * np.cos(2 * np.pi * t / (11.86 * YEAR))      # a couple of light-seconds of solar wobble
    sun[:, 1] += 1.5 * np.sin(2 * np.pi * t / (11.86 * YEAR))
    cols = [earth, vel, sun - earth]
    for _, r, per in _PLANETS:
        pos, _ = _orbit(t, r, per, rng.uniform(0, 2 * np.pi))
        cols.append(pos - earth)
    return np.concatenate(cols, axis=1)


def _records(t_hi, t_lo, err, freq, eph, wide_dm=None) -> np.ndarray:
    n = len(t_hi)
    rec = np.zeros((n, M.WTOA_WORDS if wide_dm is not None else M.TOA_WORDS))
    rec[:, 0], rec[:, 1], rec[:, 2], rec[:, 3] = t_hi, t_lo, err, freq
    rec[:, 6:30] = eph
    rec[:, 30] = np.arange(1, n + 1, dtype=np.uint64).view(np.float64)
    if wide_dm is not None:
        rec[:, 31], rec[:, 32] = wide_dm
    return rec


def _two_sum(a, b):
    s = a + b
    v = s - a
    return s, (a - (s - v)) + (b - v)


def _shift(t_hi, t_lo, delta):
    """(t_hi, t_lo) + delta in double-double (delta: float64 array)."""
    s, e = _two_sum(t_hi, delta)
    e = e + t_lo
    hi = s + e
    return hi, e - (hi - s)


def _fourier_basis(t, nharm, tspan):
    """PINT create_fourier_design_matrix convention: columns [sin x nharm, cos x nharm] (model.py:566-567)."""
    k = np.arange(1, nharm + 1)
    arg = 2 * np.pi * np.outer(t - t.min(), k) / tspan
    return np.sin(arg), np.cos(arg)


def _powerlaw_weights(log10A, gamma, f1, nharm):
    """gp_noise.jl:12-16,64-78 (weights, not inverse) for linear harmonics."""
    A = 10.0 ** log10A
    fyr = 1.0 / 3600 / 24 / 365.25
    P1 = A * A / (12 * np.pi**2 * fyr**3) * (fyr / f1) ** gamma * f1
    return P1 * np.exp(-gamma * np.log(np.arange(1, nharm + 1)))


# ---------------------------------------------------------------------------------------------
# configuration table
# ---------------------------------------------------------------------------------------------
CONFIGS = {
    # BASELINE.json configs[1]: "sim2 synthetic ELL1 binary pulsar, 5k TOAs, EFAC/EQUAD white noise, 4096-walker batch"
    "config2_ell1": dict(ntoa=5000, span_d=4000.0, binary="ELL1", nback=4, walkers=4096, seed=2),
    # configs[2]: "sim3_gp: PowerlawRedNoiseGP Woodbury likelihood, 10k TOAs, 30 Fourier harmonics, 8192 walkers"
    "config3_gp": dict(ntoa=10000, span_d=4000.0, red=30, nback=4, walkers=8192, seed=3),
    # configs[3]: "BinaryDD + DMGP + SolarWind wideband TOAs with MarginalizedTimingModel, 20k TOAs, 8192 walkers"
    "config4_dd_wb": dict(ntoa=20000, span_d=5000.0, binary="DD", wideband=True, dmgp=30, solarwind=True,
                          ecliptic=True, mtm=True, nback=1, walkers=8192, seed=4),
    # configs[4]: "100k TOAs, red+DM+chromatic GP (rank ~300), 65536 walkers sharded over 8 GPUs"
    "config5_array": dict(ntoa=100000, span_d=6000.0, red=50, dmgp=50, chromgp=50, nback=8, walkers=65536, seed=5),
    # component-coverage cases for the parity tests (not benchmark lines)
    "extra_ell1_fbx": dict(ntoa=1500, span_d=3000.0, binary="ELL1", fbx=True, glitch=2, expdip=2, dmx=6, cmx=4,
                           jump="exclusive", nback=3, walkers=256, seed=11),
    "extra_dd_fbx_wb": dict(ntoa=1200, span_d=3000.0, binary="DD", fbx=True, wideband=True, dmjump="exclusive",
                            jump="bits", solarwind=True, ecliptic=True, glitch=1, nback=2, walkers=256, seed=12),
    "extra_dmjump_bits_wb": dict(ntoa=800, span_d=2000.0, wideband=True, dmjump="bits", ecliptic=False, nback=2,
                                 red=10, walkers=128, seed=13),
    "extra_ell1h_fd_wavex": dict(ntoa=1000, span_d=2500.0, binary="ELL1H", fd=3, fdjump=2, wavex=4, dmwavex=3, cmwavex=2,
                                 nback=2, walkers=128, seed=14),
    "extra_ell1k_swx": dict(ntoa=900, span_d=2500.0, binary="ELL1k", swx=3, fdjumpdm="exclusive", ecliptic=True, nback=2,
                            walkers=128, seed=15),
    "extra_ddk_wb": dict(ntoa=800, span_d=2500.0, binary="DDK", wideband=True, fdjumpdm="bits", ecliptic=True, px=True,
                         nback=1, walkers=128, seed=16),
    "extra_ddh": dict(ntoa=700, span_d=2500.0, binary="DDH", fbx=True, nback=2, walkers=128, seed=17),
    "extra_dds": dict(ntoa=700, span_d=2500.0, binary="DDS", nback=2, walkers=128, seed=18),
    # ECORR kernels: epochs of 8 TOAs (one per observing frequency) form the ECORR groups
    "extra_ecorr": dict(ntoa=1600, span_d=3000.0, ecorr=True, nback=3, walkers=256, seed=19),
    "extra_ecorr_gp": dict(ntoa=1600, span_d=3000.0, ecorr=True, red=12, dmgp=10, nback=3, walkers=256, seed=20),
    # rank 4 (MarginalizedTimingModel alone): a single, partial block column in the Cholesky finish
    "extra_mtm_only": dict(ntoa=700, span_d=2500.0, mtm=True, nback=2, walkers=128, seed=25),
    # ranks above 64 and above 128: the wider variants of the ECORR correction kernels
    "extra_ecorr_gp_r80": dict(ntoa=1200, span_d=3000.0, ecorr=True, red=25, dmgp=15, nback=3, walkers=128, seed=23),
    "extra_ecorr_gp_r140": dict(ntoa=1200, span_d=3000.0, ecorr=True, red=40, dmgp=30, nback=2, walkers=128, seed=24),
    # power-law GPs with free amplitudes as delay components (marginalize_gp_noise=False): (Nlin, Nlog) per GP
    "extra_free_gp": dict(ntoa=900, span_d=3000.0, ured=(6, 2), udm=(5, 0), uchrom=(4, 1), nback=2, walkers=128, seed=22),
    "extra_free_gp_wb": dict(ntoa=700, span_d=2500.0, wideband=True, udm=(6, 1), ured=(4, 0), binary="ELL1", nback=2,
                             walkers=128, seed=23),
    "config3_gp_ecorr": dict(ntoa=10000, span_d=4000.0, red=30, ecorr=True, nback=4, walkers=8192, seed=21),
}


def make_dataset(name: str, make_eval: Callable, ntoa: Optional[int] = None, seed: Optional[int] = None) -> SynthDataset:
    cfg = dict(CONFIGS[name])
    if ntoa is not None:
        cfg["ntoa"] = ntoa
    if seed is not None:
        cfg["seed"] = seed
    rng = np.random.default_rng(cfg["seed"])
    n = cfg["ntoa"]
    span = cfg["span_d"] * DAY
    wide = bool(cfg.get("wideband"))
    ecl = bool(cfg.get("ecliptic"))
    nback = cfg["nback"]

    # ---- TOA skeleton
    t = np.sort(rng.uniform(-span / 2, span / 2, n))
    freqs = np.array([500e6, 643e6, 786e6, 929e6, 1071e6, 1214e6, 1357e6, 1500e6])
    freq = 1400e6 * np.ones(n) if wide else freqs[np.arange(n) % 8]
    err = rng.uniform(0.5e-6, 2e-6, n)
    backend = (np.arange(n) // 8) % nback + 1          # all 8 frequencies of an epoch share a backend
    eph = ephemeris_words(t, rng)
    dm_err = DMCONST * 1e-4 * np.ones(n)

    # ---- parameters (Vela units)
    single: List[M.Parameter] = []
    multi: List[M.MultiParameter] = []
    sigma: Dict[str, float] = {}
    noise_draw_names: Dict[str, tuple] = {}

    def sp(name, val, free=False, dim=0, noise=False, sig=0.0):
        single.append(M.Parameter(name, float(val), dim, not free, "", 1.0, noise))
        if free:
            sigma[name] = sig

    def mp(name, vals, free, dim=0, noise=False, sig=None, prefix=None):
        ps = []
        for i, (v, fr) in enumerate(zip(vals, free)):
            pname = f"{prefix or name}{i + 1}" if len(vals) > 1 or prefix else name
            ps.append(M.Parameter(pname, float(v), dim, not fr, "", 1.0, noise))
            if fr:
                sigma[pname] = sig[i] if sig is not None else 0.0
        multi.append(M.MultiParameter(name, ps))

    F0 = 218.8118437960826 if name != "config2_ell1" else 173.6879458121843
    F_hi = float(F0)
    lon, lat = (1.2, 0.3) if ecl else (4.51, -0.21)
    sp("POSEPOCH", 0.0, dim=1)
    sp("PX", 1.2 * 9.715611890180196e-12 if (ecl or cfg.get("px")) else 0.0, dim=-1)
    sp("ELONG" if ecl else "RAJ", lon, free=True)
    sp("ELAT" if ecl else "DECJ", lat, free=True)
    sp("PMELONG" if ecl else "PMRA", 3.0 * 1.5362818500441604e-16, free=True, dim=-1)
    sp("PMELAT" if ecl else "PMDEC", -2.0 * 1.5362818500441604e-16, free=True, dim=-1)
    mtm = bool(cfg.get("mtm"))
    sp("PHOFF", 0.0, free=not mtm)
    if cfg.get("solarwind"):
        sp("SWEPOCH", 0.0, dim=1)
    sp("DMEPOCH", 0.0, dim=1)
    if cfg.get("chromgp"):
        sp("CMEPOCH", 0.0, dim=1)
        sp("TNCHROMIDX", 4.0)
    free_gps = [(key, pre, tn, cfg[key]) for key, pre, tn in (("ured", "PLRED", "TNRED"), ("udm", "PLDM", "TNDM"),
                                                              ("uchrom", "PLCHROM", "TNCHROM")) if cfg.get(key)]
    for key, pre, tn, (nlin, nlog) in free_gps:
        sp(pre + "EPOCH", -0.37 * span, dim=1)
        sp(pre + "FREQ", 1.0 / span, dim=-1)
        if key == "uchrom" and not cfg.get("chromgp"):
            sp("TNCHROMIDX", 4.0)
    binary = cfg.get("binary")
    ell1_like = binary in ("ELL1", "ELL1H", "ELL1k")
    dd_like = binary in ("DD", "DDH", "DDS", "DDK")
    if ell1_like:
        sp("TASC", 0.3 * DAY, free=True, dim=1)
        if not cfg.get("fbx"):
            sp("PB", 4.0742 * DAY, free=True, dim=1)
            sp("PBDOT", 0.0)
        sp("A1", 3.3667, free=True, dim=1)
        sp("A1DOT", 0.0)
        sp("EPS1", 1.2e-5, free=True)
        sp("EPS2", -0.8e-5, free=True)
        if binary == "ELL1k":
            sp("OMDOT", 3.0e-9, free=True, dim=-1)
            sp("LNEDOT", 1.0e-12, dim=-1)
        else:
            sp("EPS1DOT", 0.0, dim=-1)
            sp("EPS2DOT", 0.0, dim=-1)
        if binary == "ELL1H":
            sp("H3", 2.0e-7, free=True, dim=1)
            sp("STIGMA", 0.55, free=True)
        else:
            sp("M2", 0.25 * 4.92549094830932e-06, free=True, dim=1)
            sp("SINI", 0.95, free=True)
    elif dd_like:
        sp("T0", 1.7 * DAY, free=True, dim=1)
        if not cfg.get("fbx"):
            sp("PB", 25.3 * DAY, free=True, dim=1)
            sp("PBDOT", 0.0)
        sp("A1", 21.3, free=True, dim=1)
        sp("A1DOT", 0.0)
        sp("ECC", 0.1072, free=True)
        sp("EDOT", 0.0, dim=-1)
        sp("OM", 1.1, free=True)
        sp("OMDOT", 2.0e-10, free=True, dim=-1)
        sp("DR", 0.0)
        sp("DTH", 0.0)
        sp("GAMMA", 1.0e-3, dim=1)
        if binary == "DDH":
            sp("H3", 3.0e-7, free=True, dim=1)
            sp("STIGMA", 0.6, free=True)
        elif binary == "DDS":
            sp("M2", 0.3 * 4.92549094830932e-06, free=True, dim=1)
            sp("SHAPMAX", 2.5, free=True)
        elif binary == "DDK":
            sp("M2", 0.3 * 4.92549094830932e-06, free=True, dim=1)
            sp("KIN", 1.1, free=True)
            sp("KOM", 0.7, free=True)
        else:
            sp("M2", 0.3 * 4.92549094830932e-06, free=True, dim=1)
            sp("SINI", 0.9, free=True)
    for key, amp, gam in (("RED", -13.6, 3.8), ("DM", -13.4, 3.2), ("CHROM", -13.8, 3.5)):
        flag = {"RED": "red", "DM": "dmgp", "CHROM": "chromgp"}[key]
        if cfg.get(flag):
            sp(f"PL{key}FREQ", 1.0 / span, dim=-1)
    sp("F_", F_hi, dim=-1)
    dm0 = 15.0 * DMCONST
    if cfg.get("solarwind"):
        mp("NE_SW", [1.7e8], [True], dim=-2, sig=[0.0])   # ~7.9 cm^-3 in Vela units (cf. sim_sw.wb fixture)
    mp("DM", [dm0, 1e-4 * DMCONST / YEAR], [not mtm, True], dim=-1, prefix="DM", sig=[0.0, 0.0])
    if cfg.get("chromgp"):
        mp("CM", [0.0], [False], dim=-1, prefix="CM")
    if cfg.get("fbx"):
        pb = (4.0742 if ell1_like else 25.3) * DAY
        mp("FB", [1.0 / pb, -2.0e-20], [True, True], dim=-1, prefix="FB", sig=[0.0, 0.0])
    ng = cfg.get("glitch", 0)
    if ng:
        gl_ep = np.linspace(-0.3, 0.25, ng) * span
        mp("GLEP_", gl_ep, [False] * ng, dim=1, prefix="GLEP_")
        mp("GLPH_", 0.1 + 0.05 * np.arange(ng), [True] * ng, prefix="GLPH_", sig=[0.0] * ng)
        mp("GLF0_", 2e-8 * (1 + np.arange(ng)), [True] * ng, dim=-1, prefix="GLF0_", sig=[0.0] * ng)
        mp("GLF1_", -1e-16 * np.ones(ng), [True] * ng, dim=-2, prefix="GLF1_", sig=[0.0] * ng)
        mp("GLF2_", np.zeros(ng), [False] * ng, dim=-3, prefix="GLF2_")
        mp("GLF0D_", 1e-8 * np.ones(ng), [True] * ng, dim=-1, prefix="GLF0D_", sig=[0.0] * ng)
        mp("GLTD_", 50.0 * DAY * np.ones(ng), [False] * ng, dim=1, prefix="GLTD_")
    ne = cfg.get("expdip", 0)
    if ne:
        sp("EXPDIPFREF", 1.4e9, dim=-1)
        sp("EXPDIPEPS", 0.1 * DAY, dim=1)
        mp("EXPDIPEP_", np.linspace(-0.2, 0.3, ne) * span, [False] * ne, dim=1, prefix="EXPDIPEP_")
        mp("EXPDIPAMP_", 2e-6 * np.ones(ne), [True] * ne, dim=1, prefix="EXPDIPAMP_", sig=[0.0] * ne)
        mp("EXPDIPIDX_", -2.0 * np.ones(ne), [True] * ne, prefix="EXPDIPIDX_", sig=[0.0] * ne)
        mp("EXPDIPTAU_", 30.0 * DAY * np.ones(ne), [True] * ne, dim=1, prefix="EXPDIPTAU_", sig=[0.0] * ne)
    ndmx = cfg.get("dmx", 0)
    if ndmx:
        mp("DMX_", 1e-3 * DMCONST * rng.standard_normal(ndmx), [True] * ndmx, dim=-1, prefix="DMX_", sig=[0.0] * ndmx)
    ncmx = cfg.get("cmx", 0)
    if ncmx:
        if not cfg.get("chromgp"):
            sp("TNCHROMIDX", 4.0)
        mp("CMX_", 1e-3 * DMCONST * rng.standard_normal(ncmx), [True] * ncmx, dim=-1, prefix="CMX_", sig=[0.0] * ncmx)
    njump = 2 if cfg.get("jump") else 0
    if njump:
        mp("JUMP", [3e-6, -2e-6], [True, True], dim=1, prefix="JUMP", sig=[0.0, 0.0])
    if cfg.get("dmjump"):
        mp("DMJUMP", [2e-4 * DMCONST, -1e-4 * DMCONST], [True, True], dim=-1, prefix="DMJUMP", sig=[0.0, 0.0])
    nfd = cfg.get("fd", 0)
    if nfd:
        mp("FD", 1e-6 * np.array([1.0, -0.4, 0.2])[:nfd], [True] * nfd, dim=1, prefix="FD", sig=[0.0] * nfd)
    nfdj = cfg.get("fdjump", 0)
    if nfdj:
        mp("FDJUMP", 5e-7 * np.array([1.0, -0.6])[:nfdj], [True] * nfdj, dim=1, prefix="FDJUMP", sig=[0.0] * nfdj)
    for pre, key, amp, dim in (("WX", "wavex", 2e-7, 1), ("DMWX", "dmwavex", 2e-4 * DMCONST, -1), ("CMWX", "cmwavex", 1e-4 * DMCONST, -1)):
        nw = cfg.get(key, 0)
        if nw:
            sp(pre + "EPOCH", 0.0, dim=1)
            if pre == "CMWX" and not (cfg.get("chromgp") or cfg.get("cmx")):
                sp("TNCHROMIDX", 4.0)
            mp(pre + "SIN_", amp * rng.standard_normal(nw), [True] * nw, dim=dim, prefix=pre + "SIN_", sig=[0.0] * nw)
            mp(pre + "COS_", amp * rng.standard_normal(nw), [True] * nw, dim=dim, prefix=pre + "COS_", sig=[0.0] * nw)
            mp(pre + "FREQ_", np.arange(1, nw + 1) / span, [False] * nw, dim=-1, prefix=pre + "FREQ_")
    nswx = cfg.get("swx", 0)
    if nswx:
        mp("SWXDM_", 3e-4 * DMCONST * (1 + 0.2 * np.arange(nswx)), [True] * nswx, dim=-1, prefix="SWXDM_", sig=[0.0] * nswx)
    if cfg.get("fdjumpdm"):
        mp("FDJUMPDM", [1e-4 * DMCONST, -2e-4 * DMCONST], [True, True], dim=-1, prefix="FDJUMPDM", sig=[0.0, 0.0])
    for key, pre, tn, (nlin, nlog) in free_gps:
        nh = nlin + nlog
        mp(pre + "SIN_", rng.standard_normal(nh), [True] * nh, prefix=pre + "SIN_", sig=[0.0] * nh)
        mp(pre + "COS_", rng.standard_normal(nh), [True] * nh, prefix=pre + "COS_", sig=[0.0] * nh)
    mp("F", [0.0, -1.0e-15], [not mtm, not mtm], dim=-1, prefix="F", sig=[0.0, 0.0])
    # noise parameters
    for (key, pre, tn, _), (amp, gam) in zip(free_gps, ((-13.6, 3.8), (-13.4, 3.2), (-13.8, 3.5))):
        amp, gam = {"ured": (-13.6, 3.8), "udm": (-13.4, 3.2), "uchrom": (-13.8, 3.5)}[key]
        sp(tn + "AMP", amp, free=True, noise=True)
        sp(tn + "GAM", gam, free=True, noise=True)
        # with fixed amplitudes the hyper-parameters scale the delay itself: keep the scatter at the few-sigma level
        noise_draw_names[tn + "AMP"] = ("normal", amp, 0.002)
        noise_draw_names[tn + "GAM"] = ("normal", gam, 0.002)
    for key, amp, gam in (("RED", -13.6, 3.8), ("DM", -13.4, 3.2), ("CHROM", -13.8, 3.5)):
        flag = {"RED": "red", "DM": "dmgp", "CHROM": "chromgp"}[key]
        if cfg.get(flag):
            sp(f"TN{key}AMP", amp, free=True, noise=True)
            sp(f"TN{key}GAM", gam, free=True, noise=True)
            noise_draw_names[f"TN{key}AMP"] = ("normal", amp, 0.3)
            noise_draw_names[f"TN{key}GAM"] = ("normal", gam, 0.4)
    efac_true = 1.0 + 0.05 * rng.standard_normal(nback)
    equad_true = np.exp(rng.uniform(np.log(1e-7), np.log(5e-7), nback))
    mp("EFAC", efac_true, [True] * nback, noise=True, prefix="EFAC", sig=[0.0] * nback)
    mp("EQUAD", equad_true, [True] * nback, dim=1, noise=True, prefix="EQUAD", sig=[0.0] * nback)
    for i in range(nback):
        noise_draw_names[f"EFAC{i + 1}"] = ("lognormal", efac_true[i], 0.05)
        noise_draw_names[f"EQUAD{i + 1}"] = ("loguniform", 0.5 * equad_true[i], 2.0 * equad_true[i])
    ecorr_true = np.exp(rng.uniform(np.log(2e-7), np.log(8e-7), nback))
    if cfg.get("ecorr"):
        mp("ECORR", ecorr_true, [True] * nback, dim=1, noise=True, prefix="ECORR", sig=[0.0] * nback)
        for i in range(nback):
            noise_draw_names[f"ECORR{i + 1}"] = ("loguniform", 0.5 * ecorr_true[i], 2.0 * ecorr_true[i])
    if wide:
        mp("DMEFAC", [1.05], [True], noise=True, prefix="DMEFAC", sig=[0.0])
        mp("DMEQUAD", [0.5e-4 * DMCONST], [True], dim=-1, noise=True, prefix="DMEQUAD", sig=[0.0])
        noise_draw_names["DMEFAC1"] = ("lognormal", 1.05, 0.05)
        noise_draw_names["DMEQUAD1"] = ("loguniform", 0.25e-4 * DMCONST, 1e-4 * DMCONST)
    handler = M.ParamHandler(single, multi)

    # ---- components in pyvela's authoritative order (model.py:85-351)
    comps: List[M.Component] = [M.SolarSystem(ecl, planet_shapiro=bool(cfg.get("solarwind")))]
    seg = lambda k: (np.minimum((np.arange(n) * (k + 1)) // n, k)).astype(np.uint32)   # 0 = none, 1..k  # noqa: E731
    if cfg.get("solarwind"):
        comps.append(M.SolarWindDispersion())
    elif nswx:
        comps.append(M.SolarWindDispersionPiecewise(seg(nswx), 5000.0, 500.0))
    comps.append(M.DispersionTaylor())
    if ndmx:
        comps.append(M.DispersionPiecewise(seg(ndmx)))
    if cfg.get("dmwavex"):
        comps.append(M.DMWaveX())
    elif cfg.get("udm"):
        comps.append(M.PowerlawDispersionNoiseGP(*cfg["udm"], 2.0))
    if cfg.get("fdjumpdm") == "exclusive":
        comps.append(M.ExclusiveDispersionOffset(((np.arange(n) // 4) % 3).astype(np.uint32)))
    elif cfg.get("fdjumpdm") == "bits":
        comps.append(M.DispersionOffset(np.stack([(np.arange(n) % 3) == 1, (np.arange(n) % 5) == 2])))
    if cfg.get("dmjump") == "exclusive":
        comps.append(M.ExclusiveDispersionJump(((np.arange(n) // 5) % 3).astype(np.uint32)))
    elif cfg.get("dmjump") == "bits":
        comps.append(M.DispersionJump(np.stack([(np.arange(n) % 3) == 0, (np.arange(n) % 4) == 1])))
    if cfg.get("chromgp"):
        comps.append(M.ChromaticTaylor())
    if ncmx:
        comps.append(M.ChromaticPiecewise(seg(ncmx)[::-1].copy()))
    if cfg.get("cmwavex"):
        comps.append(M.CMWaveX())
    elif cfg.get("uchrom"):
        comps.append(M.PowerlawChromaticNoiseGP(*cfg["uchrom"], 2.0))
    if binary == "DDK":
        comps.append(M.BinaryDDK(bool(cfg.get("fbx")), ecl))
    elif binary:
        comps.append(getattr(M, "Binary" + binary)(bool(cfg.get("fbx"))))
    if nfd:
        comps.append(M.FrequencyDependent())
    if nfdj:
        comps.append(M.FrequencyDependentJump(np.stack([(np.arange(n) % 3) == 0, (np.arange(n) % 4) == 1])[:nfdj],
                                              np.array([1, 2], dtype=np.uint32)[:nfdj]))
    if ne:
        comps.append(M.ChromaticExponentialDip())
    if cfg.get("wavex"):
        comps.append(M.WaveX())
    elif cfg.get("ured"):
        comps.append(M.PowerlawRedNoiseGP(*cfg["ured"], 2.0))
    comps.append(M.Spindown())
    if ng:
        comps.append(M.Glitch())
    comps.append(M.PhaseOffset())
    if cfg.get("jump") == "exclusive":
        comps.append(M.ExclusivePhaseJump(((np.arange(n) // 3) % 3).astype(np.uint32)))
    elif cfg.get("jump") == "bits":
        comps.append(M.PhaseJump(np.stack([(np.arange(n) % 5) < 2, (np.arange(n) % 7) == 3])))
    comps.append(M.MeasurementNoise(backend.astype(np.uint32), backend.astype(np.uint32)))
    if wide:
        ones = np.ones(n, dtype=np.uint32)
        comps.append(M.DispersionMeasurementNoise(ones, ones))

    tzr_eph = ephemeris_words(np.array([0.0]), np.random.default_rng(cfg["seed"] + 1000))[0]
    tzr = M.make_tzr_toa((0.0, 0.0), 1400e6, M.SolarSystemEphemeris(*[tuple(tzr_eph[3 * i: 3 * i + 3]) for i in range(8)]))

    def build(t_hi, t_lo, pn, kernel, dm=None):
        rec = _records(t_hi, t_lo, err, freq, eph, (dm, dm_err) if wide else None)
        rec[:, 4] = pn
        model = M.TimingModel("SYNTH_" + name, "toy-circular", "none", "TDB", 0.0, tuple(comps), kernel, handler, tzr, ())
        return model, M.TOAs(rec)

    x0 = handler.read_param_values_to_vector()

    # ---- phase connection: pulse numbers from the model phase, then Newton steps on the TOA
    t_hi, t_lo = t.copy(), np.zeros(n)
    pn = np.zeros(n)
    for it in range(4):
        model, toas = build(t_hi, t_lo, pn, M.WhiteNoiseKernel(), dm=np.zeros(n) if wide else None)
        ev = make_eval(model, toas)
        hi, lo, f = ev.phases(x0)
        if it == 0:
            pn = np.rint(hi)
            hi = hi - pn
        t_hi, t_lo = _shift(t_hi, t_lo, -(hi + lo) / f)
    model, toas = build(t_hi, t_lo, pn, M.WhiteNoiseKernel(), dm=np.zeros(n) if wide else None)
    ev = make_eval(model, toas)
    hi, lo, f = ev.phases(x0)
    connect_rms = float(np.sqrt(np.mean(((hi + lo) / f) ** 2)))

    # ---- noise realisation: white + GP, added to the TOAs; wideband DMs around the model DM
    efac_j = efac_true[backend - 1]
    equad_j = equad_true[backend - 1]
    noise = efac_j * np.sqrt(err**2 + equad_j**2) * rng.standard_normal(n)
    inner_kernel: M.Kernel = M.WhiteNoiseKernel()
    if cfg.get("ecorr"):
        # one ECORR group per epoch of 8 TOAs; the first 16 TOAs carry no ECORR (index 0), as pyvela's ecorr_sort
        # puts un-grouped TOAs first (pyvela/pyvela/ecorr.py:52-57)
        starts = list(range(16, n, 8))
        groups = [(1, 16, 0)] + [(a + 1, min(a + 8, n), int(backend[a])) for a in starts]
        inner_kernel = M.EcorrKernel(np.array(groups, dtype=np.uint64))
        for a0, a1, idx in groups:
            if idx:
                noise[a0 - 1:a1] += ecorr_true[idx - 1] * rng.standard_normal()
    tspan = t.max() - t.min()
    nu_ratio2 = (1400e6 / freq) ** 2
    blocks, gps = [], []
    toa_rows_extra: Dict[str, np.ndarray] = {}
    for key, cls, amp, gam, scale in (("red", M.PowerlawRedNoiseGP, -13.6, 3.8, np.ones(n)),
                                      ("dmgp", M.PowerlawDispersionNoiseGP, -13.4, 3.2, nu_ratio2),
                                      ("chromgp", M.PowerlawChromaticNoiseGP, -13.8, 3.5, nu_ratio2**2)):
        nh = cfg.get(key)
        if not nh:
            continue
        s, c = _fourier_basis(t, nh, tspan)
        s, c = s * scale[:, None], c * scale[:, None]
        wts = _powerlaw_weights(amp, gam, 1.0 / span, nh)
        coef = rng.standard_normal(2 * nh) * np.sqrt(np.concatenate([wts, wts]))
        noise = noise + np.concatenate([s, c], axis=1) @ coef
        if wide:   # DM-row block: toa block x nu^2 for the DM GP, zero otherwise (model.py:556-564)
            fac = freq[:, None] ** 2 if key == "dmgp" else 0.0
            s, c = np.vstack([s, s * fac]), np.vstack([c, c * fac])
            if key == "dmgp":
                toa_rows_extra["dm_gp"] = (np.concatenate([s[n:], c[n:]], axis=1) @ coef)
        blocks += [s, c]
        gps.append(cls(nh, 0, 2.0))
    t_hi, t_lo = _shift(t_hi, t_lo, noise)

    dm_meas = None
    if wide:
        model, toas = build(t_hi, t_lo, pn, M.WhiteNoiseKernel(), dm=np.zeros(n))
        y, _ = make_eval(model, toas).resids_and_Ninvdiag(x0)
        model_dm = -y[n:]
        dm_sig = 1.05 * np.sqrt(dm_err**2 + (0.5e-4 * DMCONST) ** 2)
        dm_meas = model_dm + toa_rows_extra.get("dm_gp", 0.0) + dm_sig * rng.standard_normal(n)

    if mtm:   # linearised timing parameters PHOFF, F0, F1, DM as design-matrix columns (model.py:585-619)
        cols = [np.full(n, 1.0 / F0), -t / F0, -0.5 * t**2 / F0, 1.0 / freq**2]
        dmrows = [np.zeros(n), np.zeros(n), np.zeros(n), np.ones(n)]
        blocks.append(np.stack([np.concatenate([c, d]) if wide else c for c, d in zip(cols, dmrows)], axis=1))
        gps.append(M.MarginalizedTimingModel(weights=np.full(4, 1e40), pnames=["PHOFF", "F0", "F1", "DM"]))

    kernel = M.WoodburyKernel(inner_kernel, gps, np.concatenate(blocks, axis=1)) if gps else inner_kernel
    model, toas = build(t_hi, t_lo, pn, kernel, dm=dm_meas)
    ev = make_eval(model, toas)

    # ---- walker scatter: diagonal Fisher widths from finite-difference residual derivatives
    names = handler.get_free_param_names()
    y0, w0 = ev.resids_and_Ninvdiag(x0)
    sig = np.zeros(len(names))
    noise_draw: Dict[int, tuple] = {}
    for k, nme in enumerate(names):
        if nme in noise_draw_names:
            noise_draw[k] = noise_draw_names[nme]
            continue
        # geometric search for a step that moves the residuals by ~1 sigma in total (chi ~ 1)
        h = max(abs(x0[k]) * 1e-9, 1e-30)
        chi = 0.0
        for _ in range(60):
            xp = x0.copy()
            xp[k] += h
            yk, _ = ev.resids_and_Ninvdiag(xp)
            d = (yk - y0)
            chi = float(np.sqrt(np.sum(d * d * w0)))
            if np.isfinite(chi) and 0.05 < chi < 20.0:
                break
            if not np.isfinite(chi):
                h /= 1024.0
            elif chi < 1e-3:          # below the rounding noise of the residuals: no usable slope information yet
                h *= 1024.0
            else:
                h *= min(max(1.0 / chi, 1.0 / 1024), 1024.0)
        sig[k] = h / chi if (np.isfinite(chi) and 0.05 < chi < 20.0) else 0.0
    info = dict(config=name, ntoa=n, wideband=wide, nfree=len(names), rank=int(kernel.noise_basis.shape[1]) if gps else 0,
                connect_rms_s=connect_rms, free_names=names)
    return SynthDataset(name, model, toas, x0, sig, noise_draw, info)
