"""Flatten a (TimingModel, TOAs) pair into the C-ABI `VelaModelDesc` of include/vela_b200.h.

This is the Python twin of the Julia shim's `_describe(model, toas)` (vela.jl_b200/julia/VelaB200.jl):
it walks the ordered component tuple, looks every named parameter up in the ParamHandler's
flattened ordering (src/parameter/parameter.jl:87-92) and records 0-based offsets.
"""

from __future__ import annotations

import ctypes as C
from typing import Any, List

import numpy as np

from . import model as M

VELA_B200_ABI_VERSION = 3
VELA_MAX_PAR = 24

# VelaComponentKind
(COMP_SOLAR_SYSTEM, COMP_SOLAR_WIND, COMP_DISPERSION_TAYLOR, COMP_DISPERSION_PIECEWISE,
 COMP_CHROMATIC_TAYLOR, COMP_CHROMATIC_PIECEWISE, COMP_BINARY_ELL1, COMP_BINARY_DD, COMP_GLITCH,
 COMP_EXP_DIP, COMP_SPINDOWN, COMP_PHASE_OFFSET, COMP_PHASE_JUMP_EXCLUSIVE, COMP_PHASE_JUMP,
 COMP_MEASUREMENT_NOISE, COMP_DM_JUMP_EXCLUSIVE, COMP_DM_JUMP, COMP_DM_MEASUREMENT_NOISE, COMP_BINARY_ELL1H,
 COMP_BINARY_ELL1K, COMP_BINARY_DDH, COMP_BINARY_DDS, COMP_BINARY_DDK, COMP_FD, COMP_FD_JUMP, COMP_WAVEX, COMP_DM_WAVEX,
 COMP_CM_WAVEX, COMP_SOLAR_WIND_PIECEWISE, COMP_DM_OFFSET_EXCLUSIVE, COMP_DM_OFFSET, COMP_POWERLAW_RED_GP,
 COMP_POWERLAW_DM_GP, COMP_POWERLAW_CHROM_GP) = range(1, 35)
FLAG_ECLIPTIC, FLAG_PLANET_SHAPIRO, FLAG_USE_FBX = 1, 2, 4
KERNEL_WHITE, KERNEL_ECORR, KERNEL_WOODBURY_WHITE, KERNEL_WOODBURY_ECORR = 0, 1, 2, 3
GP_POWERLAW_RED, GP_POWERLAW_DM, GP_POWERLAW_CHROM, GP_MARGINALIZED_TM = 1, 2, 3, 4


class VelaComponentDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("flags", C.c_int32), ("par", C.c_int32 * VELA_MAX_PAR),
                ("len", C.c_int32 * 4), ("mask", C.c_int32 * 2), ("dval", C.c_double * 2)]


class VelaGPDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n", C.c_int32), ("par_log10_amp", C.c_int32),
                ("par_gamma", C.c_int32), ("par_f1", C.c_int32), ("_pad", C.c_int32),
                ("ln_js", C.POINTER(C.c_double)), ("weights_inv", C.POINTER(C.c_double))]


class VelaModelDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("n_components", C.c_int32),
        ("components", C.POINTER(VelaComponentDesc)),
        ("n_params", C.c_int32), ("n_free", C.c_int32),
        ("default_values", C.POINTER(C.c_double)), ("free_indices", C.POINTER(C.c_int32)),
        ("n_toas", C.c_int64), ("wideband", C.c_int32), ("toa_stride_bytes", C.c_int32),
        ("toas", C.c_void_p), ("tzr_toa", C.c_void_p),
        ("n_index_masks", C.c_int32), ("n_bit_mask_rows", C.c_int32),
        ("index_masks", C.POINTER(C.c_uint32)), ("bit_masks", C.POINTER(C.c_uint8)),
        ("kernel_kind", C.c_int32), ("n_gp", C.c_int32), ("gp", C.POINTER(VelaGPDesc)),
        ("basis_rows", C.c_int64), ("basis_cols", C.c_int64), ("basis_ld", C.c_int64),
        ("noise_basis", C.POINTER(C.c_double)),
        ("n_ecorr_groups", C.c_int32), ("par_ecorr", C.c_int32), ("ecorr_groups", C.POINTER(C.c_uint64)),
        ("n_devices", C.c_int32), ("_pad", C.c_int32), ("devices", C.POINTER(C.c_int32)),
        ("n_aux", C.c_int64), ("aux_values", C.POINTER(C.c_double)),
    ]


class VelaPriorDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("a", C.c_double), ("b", C.c_double),
                ("lo", C.c_double), ("hi", C.c_double)]


(PRIOR_UNIFORM, PRIOR_NORMAL, PRIOR_LOGNORMAL, PRIOR_LOGUNIFORM, PRIOR_KIN, PRIOR_SINI, PRIOR_STIGMA, PRIOR_SHAPMAX,
 PRIOR_PX, PRIOR_LATITUDE, PRIOR_TRUNCATED_NORMAL) = range(1, 12)


def prior_descriptors(priors):
    """One `VelaPriorDesc` per free parameter from the host-side prior objects, or None if a family has no device twin."""
    from . import priors as P
    out = (VelaPriorDesc * max(len(priors), 1))()
    for k, pr in enumerate(priors):
        d = pr.distribution
        e = out[k]
        if isinstance(d, P.Uniform):
            e.kind, e.a, e.b = PRIOR_UNIFORM, d.a, d.b
        elif isinstance(d, P.Normal):
            e.kind, e.a, e.b = PRIOR_NORMAL, d.mu, d.sigma
        elif isinstance(d, P.LogNormal):
            e.kind, e.a, e.b = PRIOR_LOGNORMAL, d.mu, d.sigma
        elif isinstance(d, P.LogUniform):
            e.kind, e.a, e.b = PRIOR_LOGUNIFORM, d.a, d.b
        elif isinstance(d, P.TruncatedNormal):
            e.kind, e.a, e.b, e.lo, e.hi = PRIOR_TRUNCATED_NORMAL, d.mu, d.sigma, d.lower, d.upper
        elif isinstance(d, P.PXPriorDistribution):
            e.kind, e.a = PRIOR_PX, d.Rmax
        elif isinstance(d, P.KINPriorDistribution):
            e.kind = PRIOR_KIN
        elif isinstance(d, P.SINIPriorDistribution):
            e.kind = PRIOR_SINI
        elif isinstance(d, P.STIGMAPriorDistribution):
            e.kind = PRIOR_STIGMA
        elif isinstance(d, P.SHAPMAXPriorDistribution):
            e.kind = PRIOR_SHAPMAX
        elif isinstance(d, P.LatitudePriorDistribution):
            e.kind = PRIOR_LATITUDE
        else:
            return None
    return out


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class ModelDescriptor:
    """Owns the numpy buffers a `VelaModelDesc` points into."""

    def __init__(self, model: M.TimingModel, toas: M.TOAs, devices=None):
        self.model = model
        self.toas = toas
        self._keep: List[Any] = []
        ph = model.param_handler
        ntoa = len(toas)
        if ntoa == 0:
            raise ValueError("empty TOA set")

        index_masks: List[np.ndarray] = []
        bit_rows: List[np.ndarray] = []

        def add_index_mask(mask) -> int:
            mask = np.asarray(mask).astype(np.uint32)
            if mask.shape != (ntoa,):
                raise ValueError(f"index mask has shape {mask.shape}, expected ({ntoa},)")
            index_masks.append(mask)
            return len(index_masks) - 1

        def add_bit_mask(mask, n_amplitudes=None) -> int:
            mask = np.asarray(mask).astype(np.uint8)
            if mask.ndim != 2 or mask.shape[1] != ntoa:
                raise ValueError(f"bit mask has shape {mask.shape}, expected (njump, {ntoa})")
            # the reference contracts the amplitude vector with a mask column (jump.jl:14-22: dot of equal lengths)
            if n_amplitudes is not None and mask.shape[0] != n_amplitudes:
                raise ValueError(f"bit mask has {mask.shape[0]} rows for {n_amplitudes} amplitudes")
            first = len(bit_rows)
            bit_rows.extend(list(mask))
            return first

        comps: List[VelaComponentDesc] = []
        aux: List[float] = []

        def new(kind, flags=0):
            c = VelaComponentDesc()
            c.kind, c.flags = kind, flags
            for i in range(VELA_MAX_PAR):
                c.par[i] = -1
            c.mask[0] = c.mask[1] = -1
            return c

        def setp(c, slot, name, len_slot=None, required=True):
            if not ph.has(name):
                if required:
                    raise KeyError(f"component needs parameter {name}, absent from the ParamHandler")
                return False
            c.par[slot] = ph.offset(name)
            if len_slot is not None:
                c.len[len_slot] = ph.length(name)
            return True

        for comp in model.components:
            if isinstance(comp, M.SolarSystem):
                ecl = comp.ecliptic_coordinates
                c = new(COMP_SOLAR_SYSTEM, (FLAG_ECLIPTIC if ecl else 0) |
                        (FLAG_PLANET_SHAPIRO if comp.planet_shapiro else 0))
                names = ("ELONG", "ELAT", "PMELONG", "PMELAT") if ecl else ("RAJ", "DECJ", "PMRA", "PMDEC")
                for s, nme in enumerate(names + ("PX", "POSEPOCH")):
                    setp(c, s, nme)
            elif isinstance(comp, M.SolarWindDispersion):
                c = new(COMP_SOLAR_WIND)
                setp(c, 0, "SWEPOCH")
                setp(c, 1, "NE_SW", 0)
            elif isinstance(comp, M.DispersionTaylor):
                c = new(COMP_DISPERSION_TAYLOR)
                setp(c, 0, "DMEPOCH")
                setp(c, 1, "DM", 0)
            elif isinstance(comp, M.DispersionPiecewise):
                c = new(COMP_DISPERSION_PIECEWISE)
                setp(c, 0, "DMX_", 0)
                c.mask[0] = add_index_mask(comp.dmx_mask)
            elif isinstance(comp, M.ChromaticTaylor):
                c = new(COMP_CHROMATIC_TAYLOR)
                setp(c, 0, "CMEPOCH")
                setp(c, 1, "CM", 0)
                setp(c, 2, "TNCHROMIDX")
            elif isinstance(comp, M.ChromaticPiecewise):
                c = new(COMP_CHROMATIC_PIECEWISE)
                setp(c, 0, "CMX_", 0)
                setp(c, 2, "TNCHROMIDX")
                c.mask[0] = add_index_mask(comp.cmx_mask)
            elif isinstance(comp, (M.BinaryELL1, M.BinaryELL1H, M.BinaryELL1k)):
                kind = {M.BinaryELL1: COMP_BINARY_ELL1, M.BinaryELL1H: COMP_BINARY_ELL1H,
                        M.BinaryELL1k: COMP_BINARY_ELL1K}[type(comp)]
                c = new(kind, FLAG_USE_FBX if comp.use_fbx else 0)
                setp(c, 0, "TASC")
                if comp.use_fbx:
                    setp(c, 3, "FB", 0)
                else:
                    setp(c, 1, "PB")
                    setp(c, 2, "PBDOT")
                rates = ("OMDOT", "LNEDOT") if kind == COMP_BINARY_ELL1K else ("EPS1DOT", "EPS2DOT")
                shap = ("H3", "STIGMA") if kind == COMP_BINARY_ELL1H else ("M2", "SINI")
                for s, nme in enumerate(("A1", "A1DOT", "EPS1", "EPS2") + rates + shap, 4):
                    setp(c, s, nme)
            elif isinstance(comp, (M.BinaryDD, M.BinaryDDH, M.BinaryDDS, M.BinaryDDK)):
                kind = {M.BinaryDD: COMP_BINARY_DD, M.BinaryDDH: COMP_BINARY_DDH, M.BinaryDDS: COMP_BINARY_DDS,
                        M.BinaryDDK: COMP_BINARY_DDK}[type(comp)]
                flags = FLAG_USE_FBX if comp.use_fbx else 0
                if kind == COMP_BINARY_DDK and comp.ecliptic_coordinates:
                    flags |= FLAG_ECLIPTIC
                c = new(kind, flags)
                setp(c, 0, "T0")
                if comp.use_fbx:
                    setp(c, 3, "FB", 0)
                else:
                    setp(c, 1, "PB")
                    setp(c, 2, "PBDOT")
                shap = {COMP_BINARY_DD: ("M2", "SINI"), COMP_BINARY_DDH: ("H3", "STIGMA"), COMP_BINARY_DDS: ("M2", "SHAPMAX"),
                        COMP_BINARY_DDK: ("M2", "KIN", "KOM") + (("PMELONG", "PMELAT") if flags & FLAG_ECLIPTIC
                                                               else ("PMRA", "PMDEC")) + ("PX",)}[kind]
                for s, nme in enumerate(("A1", "A1DOT", "ECC", "EDOT", "OM", "OMDOT", "DR", "DTH", "GAMMA") + shap, 4):
                    setp(c, s, nme)
            elif isinstance(comp, M.FrequencyDependent):
                c = new(COMP_FD)
                setp(c, 0, "FD", 0)
            elif isinstance(comp, M.FrequencyDependentJump):
                c = new(COMP_FD_JUMP)
                setp(c, 0, "FDJUMP", 0)
                c.mask[0] = add_bit_mask(comp.jump_mask)
                exps = np.zeros(ntoa, dtype=np.uint32)
                if len(comp.exponents) > ntoa or len(comp.exponents) != np.asarray(comp.jump_mask).shape[0]:
                    raise ValueError("FDJUMP exponents must match the mask rows")
                exps[: len(comp.exponents)] = comp.exponents
                c.mask[1] = add_index_mask(exps)
            elif isinstance(comp, (M.WaveX, M.DMWaveX, M.CMWaveX)):
                pre, kind = {M.WaveX: ("WX", COMP_WAVEX), M.DMWaveX: ("DMWX", COMP_DM_WAVEX),
                             M.CMWaveX: ("CMWX", COMP_CM_WAVEX)}[type(comp)]
                c = new(kind)
                setp(c, 0, pre + "EPOCH")
                setp(c, 1, pre + "SIN_", 0)
                setp(c, 2, pre + "COS_")
                setp(c, 3, pre + "FREQ_")
                if kind == COMP_CM_WAVEX:
                    setp(c, 4, "TNCHROMIDX")
            elif isinstance(comp, M._PowerlawGP):
                # the same structs double as delay components when their amplitudes are free parameters
                kind = {M.PowerlawRedNoiseGP: COMP_POWERLAW_RED_GP, M.PowerlawDispersionNoiseGP: COMP_POWERLAW_DM_GP,
                        M.PowerlawChromaticNoiseGP: COMP_POWERLAW_CHROM_GP}[type(comp)]
                c = new(kind)
                setp(c, 0, comp.amp)
                setp(c, 1, comp.gam)
                setp(c, 2, comp.prefix + "SIN_", 0)
                setp(c, 3, comp.prefix + "COS_")
                setp(c, 4, comp.freq)
                setp(c, 5, comp.prefix + "EPOCH")
                if kind == COMP_POWERLAW_CHROM_GP:
                    setp(c, 6, "TNCHROMIDX")
                if c.len[0] != len(comp.ln_js) or ph.length(comp.prefix + "COS_") != len(comp.ln_js):
                    raise ValueError("length(alphas) == length(betas) == length(ln_js) (gp_noise.jl:24)")
                zeros = np.flatnonzero(comp.ln_js == 0)
                if len(zeros) == 0:
                    raise ValueError("ln_js has no zero entry (findfirst(iszero, ln_js), gp_noise.jl:37)")
                c.len[1] = int(zeros[0])
                c.len[2] = len(aux)
                aux.extend(comp.ln_js.tolist())
                aux.extend(comp.js().tolist())
            elif isinstance(comp, M.SolarWindDispersionPiecewise):
                c = new(COMP_SOLAR_WIND_PIECEWISE)
                setp(c, 0, "SWXDM_", 0)
                c.mask[0] = add_index_mask(comp.swx_mask)
                c.dval[0], c.dval[1] = comp.conjunction_geometry, comp.opposition_geometry
            elif isinstance(comp, (M.ExclusiveDispersionOffset, M.DispersionOffset)):
                excl = isinstance(comp, M.ExclusiveDispersionOffset)
                c = new(COMP_DM_OFFSET_EXCLUSIVE if excl else COMP_DM_OFFSET)
                setp(c, 0, "FDJUMPDM", 0)
                c.mask[0] = add_index_mask(comp.jump_mask) if excl else add_bit_mask(comp.jump_mask, c.len[0])
            elif isinstance(comp, M.Glitch):
                c = new(COMP_GLITCH)
                for s, nme in enumerate(("GLEP_", "GLPH_", "GLF0_", "GLF1_", "GLF2_", "GLF0D_", "GLTD_")):
                    setp(c, s, nme, 0)
            elif isinstance(comp, M.ChromaticExponentialDip):
                c = new(COMP_EXP_DIP)
                setp(c, 0, "EXPDIPFREF")
                setp(c, 1, "EXPDIPEPS")
                for s, nme in enumerate(("EXPDIPEP_", "EXPDIPAMP_", "EXPDIPIDX_", "EXPDIPTAU_"), 2):
                    setp(c, s, nme, 0)
            elif isinstance(comp, M.Spindown):
                c = new(COMP_SPINDOWN)
                setp(c, 0, "F_")
                setp(c, 1, "F", 0)
            elif isinstance(comp, M.PhaseOffset):
                c = new(COMP_PHASE_OFFSET)
                setp(c, 0, "PHOFF")
            elif isinstance(comp, (M.ExclusivePhaseJump, M.PhaseJump)):
                excl = isinstance(comp, M.ExclusivePhaseJump)
                c = new(COMP_PHASE_JUMP_EXCLUSIVE if excl else COMP_PHASE_JUMP)
                setp(c, 0, "JUMP", 0)
                setp(c, 1, "F_")
                setp(c, 2, "F")
                c.mask[0] = add_index_mask(comp.jump_mask) if excl else add_bit_mask(comp.jump_mask, c.len[0])
            elif isinstance(comp, M.MeasurementNoise):
                c = new(COMP_MEASUREMENT_NOISE)
                setp(c, 0, "EFAC", 0, required=False)
                setp(c, 1, "EQUAD", 1, required=False)
                c.mask[0] = add_index_mask(comp.efac_index_mask)
                c.mask[1] = add_index_mask(comp.equad_index_mask)
            elif isinstance(comp, (M.ExclusiveDispersionJump, M.DispersionJump)):
                excl = isinstance(comp, M.ExclusiveDispersionJump)
                c = new(COMP_DM_JUMP_EXCLUSIVE if excl else COMP_DM_JUMP)
                setp(c, 0, "DMJUMP", 0)
                c.mask[0] = add_index_mask(comp.jump_mask) if excl else add_bit_mask(comp.jump_mask, c.len[0])
            elif isinstance(comp, M.DispersionMeasurementNoise):
                c = new(COMP_DM_MEASUREMENT_NOISE)
                setp(c, 0, "DMEFAC", 0, required=False)
                setp(c, 1, "DMEQUAD", 1, required=False)
                c.mask[0] = add_index_mask(comp.dmefac_index_mask)
                c.mask[1] = add_index_mask(comp.dmequad_index_mask)
            else:
                raise NotImplementedError(f"component {type(comp).__name__} is outside the implemented hot path")
            comps.append(c)

        d = VelaModelDesc()
        d.abi_version = VELA_B200_ABI_VERSION
        self._comps = (VelaComponentDesc * len(comps))(*comps)
        d.n_components = len(comps)
        d.components = self._comps

        self._aux = np.ascontiguousarray(aux, dtype=np.float64)
        d.n_aux = len(self._aux)
        if len(self._aux):
            d.aux_values = _dptr(self._aux)
        self._defaults = np.ascontiguousarray(ph._default_values, dtype=np.float64)
        self._free = np.ascontiguousarray(ph._free_indices0, dtype=np.int32)
        d.n_params = len(self._defaults)
        d.n_free = len(self._free)
        d.default_values = _dptr(self._defaults)
        d.free_indices = self._free.ctypes.data_as(C.POINTER(C.c_int32))

        self._records = np.ascontiguousarray(toas.records, dtype=np.float64)
        self._tzr = np.ascontiguousarray(model.tzr_toa.record(), dtype=np.float64)
        d.n_toas = ntoa
        d.wideband = 1 if toas.wideband else 0
        d.toa_stride_bytes = self._records.shape[1] * 8
        d.toas = self._records.ctypes.data
        d.tzr_toa = self._tzr.ctypes.data

        self._index_masks = (np.ascontiguousarray(np.stack(index_masks), dtype=np.uint32)
                             if index_masks else np.zeros((0, ntoa), dtype=np.uint32))
        self._bit_masks = (np.ascontiguousarray(np.stack(bit_rows), dtype=np.uint8)
                           if bit_rows else np.zeros((0, ntoa), dtype=np.uint8))
        d.n_index_masks = self._index_masks.shape[0]
        d.n_bit_mask_rows = self._bit_masks.shape[0]
        d.index_masks = self._index_masks.ctypes.data_as(C.POINTER(C.c_uint32))
        d.bit_masks = self._bit_masks.ctypes.data_as(C.POINTER(C.c_uint8))

        kernel = model.kernel
        inner = kernel.inner_kernel if isinstance(kernel, M.WoodburyKernel) else kernel
        d.n_ecorr_groups, d.par_ecorr = 0, -1
        if isinstance(inner, M.EcorrKernel):
            if toas.wideband:
                raise ValueError("ECORR kernel is not supported for wideband TOAs (pyvela/pyvela/model.py:470)")
            grp = np.ascontiguousarray(inner.ecorr_groups, dtype=np.uint64).reshape(-1, 3)
            if len(grp) == 0 or grp[0, 0] != 1 or grp[-1, 1] != ntoa or np.any(grp[1:, 0] != grp[:-1, 1] + 1):
                raise ValueError("ECORR groups must partition the TOAs into contiguous 1-based ranges")
            self._ecorr_groups = grp
            d.n_ecorr_groups = len(grp)
            d.ecorr_groups = grp.ctypes.data_as(C.POINTER(C.c_uint64))
            if grp[:, 2].max() > 0:
                d.par_ecorr = ph.offset("ECORR")
                if grp[:, 2].max() > ph.length("ECORR"):
                    raise ValueError("ECORR group index exceeds the number of ECORR parameters")
        elif not isinstance(inner, M.WhiteNoiseKernel):
            raise NotImplementedError(f"kernel {kernel!r} is outside the implemented hot path")
        ecorr = isinstance(inner, M.EcorrKernel)
        if not isinstance(kernel, M.WoodburyKernel):
            d.kernel_kind = KERNEL_ECORR if ecorr else KERNEL_WHITE
            d.n_gp = 0
        else:
            d.kernel_kind = KERNEL_WOODBURY_ECORR if ecorr else KERNEL_WOODBURY_WHITE
            gps = []
            for gp in kernel.gp_components:
                g = VelaGPDesc()
                g.par_log10_amp = g.par_gamma = g.par_f1 = -1
                if isinstance(gp, M.MarginalizedTimingModel):
                    g.kind = GP_MARGINALIZED_TM
                    w = np.ascontiguousarray(gp.prior_weights_inv, dtype=np.float64)
                    self._keep.append(w)
                    g.n = len(w)
                    g.weights_inv = _dptr(w)
                else:
                    g.kind = {M.PowerlawRedNoiseGP: GP_POWERLAW_RED, M.PowerlawDispersionNoiseGP: GP_POWERLAW_DM,
                              M.PowerlawChromaticNoiseGP: GP_POWERLAW_CHROM}[type(gp)]
                    lj = np.ascontiguousarray(gp.ln_js, dtype=np.float64)
                    self._keep.append(lj)
                    g.n = len(lj)
                    g.ln_js = _dptr(lj)
                    g.par_log10_amp = ph.offset(gp.amp)
                    g.par_gamma = ph.offset(gp.gam)
                    g.par_f1 = ph.offset(gp.freq)
                gps.append(g)
            self._gps = (VelaGPDesc * len(gps))(*gps)
            d.n_gp = len(gps)
            d.gp = self._gps
            self._basis = np.asfortranarray(kernel.noise_basis, dtype=np.float64)
            nrows = 2 * ntoa if toas.wideband else ntoa
            if self._basis.shape[0] != nrows:
                raise ValueError(f"noise_basis has {self._basis.shape[0]} rows, expected {nrows}")
            d.basis_rows, d.basis_cols = self._basis.shape
            d.basis_ld = self._basis.shape[0]
            d.noise_basis = _dptr(self._basis)

        if devices:
            self._devices = np.ascontiguousarray(devices, dtype=np.int32)
            d.n_devices = len(self._devices)
            d.devices = self._devices.ctypes.data_as(C.POINTER(C.c_int32))
        else:
            d.n_devices = 0
        self.desc = d

    @property
    def n_rows(self) -> int:
        return (2 if self.toas.wideband else 1) * len(self.toas)

    @property
    def n_basis(self) -> int:
        return int(self.desc.basis_cols) if self.desc.kernel_kind in (KERNEL_WOODBURY_WHITE, KERNEL_WOODBURY_ECORR) else 0

    def byref(self):
        return C.byref(self.desc)
