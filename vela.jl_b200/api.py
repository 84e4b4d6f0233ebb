"""The drop-in surface: Vela.jl's public names for the posterior hot path, backed by libvela_b200.so.

Reference names mirrored (paths relative to the reference tree):
  Pulsar                                      src/pulsar/pulsar.jl:9-25
  calc_lnlike / calc_lnlike_serial            src/likelihood/wls_likelihood.jl:24-56, gls_likelihood.jl:249-296
  calc_lnlike_vectorized                      src/likelihood/gls_likelihood.jl:298-309
  calc_lnprior / prior_transform              src/prior/prior.jl:13-37
  calc_lnpost / _serial / _vectorized         src/likelihood/posterior.jl:4-25
  get_lnlike_{serial,parallel}_func, get_lnlike_func   src/likelihood/wls_likelihood.jl:75-104
  get_lnpost_func                             src/likelihood/posterior.jl:34-46
  get_lnprior_func / get_prior_transform_func src/prior/prior.jl:26,37
  calc_chi2(_serial), get_chi2_*func, degrees_of_freedom, calc_chi2_reduced   src/likelihood/wls_chi2.jl
  form_residuals / _calc_resids_and_Ninvdiag  src/residuals/residuals.jl:19-26, gls_likelihood.jl:9-79
  calc_noise_weights_inv                      src/likelihood/gls_likelihood.jl:3-7
Every entry point accepts either (model, toas, params) or (pulsar, params), and `params` may be the
free-value vector or the full parameter vector returned by `read_params` (the NamedTuple in Julia).
Serial, parallel and vectorised flavours all map onto the same GPU kernels (B = 1 is a TOA-parallel
launch); there is no CPU path.
"""

from __future__ import annotations

import ctypes as C
import weakref
from typing import Callable, Optional

import numpy as np

from . import lib as _lib
from . import priors as _priors
from .descriptor import ModelDescriptor, prior_descriptors
from .model import TimingModel, TOAs, WhiteNoiseKernel


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Pulsar:
    """`Pulsar(model, toas)`: the object `pyvela.SPNTA` holds; owns the GPU-resident replica."""

    def __init__(self, model: TimingModel, toas: TOAs, devices=None):
        self.model = model
        self.toas = toas
        self.descriptor = ModelDescriptor(model, toas, devices=devices)
        self._lib = _lib.lib()
        if self._lib.vela_b200_device_count() <= 0:
            raise _lib.VelaB200Error("no CUDA device: vela.jl_b200 has no CPU fallback")
        h = C.c_void_p()
        _lib.check(self._lib.vela_b200_create(self.descriptor.byref(), C.byref(h)), "vela_b200_create")
        self._h = h
        self._finalizer = weakref.finalize(self, self._lib.vela_b200_destroy, h)
        self.ndim = model.param_handler._nfree
        self.n_rows = int(self._lib.vela_b200_n_rows(h))
        self.device_priors = False
        self.set_priors(model.priors)

    def set_priors(self, priors) -> bool:
        """Upload the per-parameter priors to the device when every family has a device twin (§8f rank 3);
        otherwise lnprior stays on the host, exactly as in the reference (posterior.jl:9-12)."""
        self.model.priors = tuple(priors)
        self.device_priors = False
        if len(self.model.priors) == self.ndim and self.ndim > 0:
            descs = prior_descriptors(self.model.priors)
            if descs is not None:
                _lib.check(self._lib.vela_b200_set_priors(self._h, descs, self.ndim), "vela_b200_set_priors")
                self.device_priors = True
        return self.device_priors

    # -- helpers
    def _free(self, params) -> np.ndarray:
        """Accept the free vector or the full parameter vector (NamedTuple analogue)."""
        p = np.ascontiguousarray(params, dtype=np.float64)
        ph = self.model.param_handler
        if p.ndim == 1 and p.shape[0] == len(ph._default_values) and p.shape[0] != ph._nfree:
            p = ph.read_param_values_to_vector(p)
        return p

    def _batch(self, fn, paramss, *extra) -> np.ndarray:
        a = np.asarray(paramss, dtype=np.float64)
        if a.ndim != 2:
            raise ValueError("paramss must be a (nsamples, ndim) matrix")
        B, nd = a.shape
        out = np.empty(B, dtype=np.float64)
        rs, cs = (s // 8 for s in a.strides) if B > 0 and nd > 0 else (nd, 1)
        if B > 0 and (a.strides[0] % 8 or a.strides[1] % 8 or rs < 0 or cs < 0):
            a = np.ascontiguousarray(a)
            rs, cs = nd, 1
        _lib.check(fn(self._h, _dp(a), B, nd, rs, cs, *extra, _dp(out)), fn.__name__)
        return out

    # -- likelihood
    def lnlike_vectorized(self, paramss) -> np.ndarray:
        return self._batch(self._lib.vela_b200_lnlike_batch, paramss)

    def lnlike(self, params) -> float:
        return float(self.lnlike_vectorized(self._free(params)[None, :])[0])

    def chi2_vectorized(self, paramss) -> np.ndarray:
        return self._batch(self._lib.vela_b200_chi2_batch, paramss)

    def chi2(self, params) -> float:
        return float(self.chi2_vectorized(self._free(params)[None, :])[0])

    # -- prior / posterior
    def lnprior(self, params):
        return _priors.lnprior(self.model.priors, self._free(params) if np.ndim(params) == 1 else params)

    def prior_transform(self, cube):
        return _priors.prior_transform(self.model.priors, cube)

    def lnprior_vectorized(self, paramss) -> np.ndarray:
        a = np.asarray(paramss, dtype=np.float64)
        if self.device_priors:
            return self._batch(self._lib.vela_b200_lnprior_batch, a)
        return np.asarray(_priors.lnprior(self.model.priors, a), dtype=np.float64)

    def prior_transform_vectorized(self, cubes) -> np.ndarray:
        q = np.ascontiguousarray(np.atleast_2d(cubes), dtype=np.float64)
        if not self.device_priors:
            return _priors.prior_transform(self.model.priors, q)
        out = np.empty_like(q)
        _lib.check(self._lib.vela_b200_prior_transform_batch(self._h, _dp(q), q.shape[0], q.shape[1], _dp(out)),
                   "vela_b200_prior_transform_batch")
        return out

    def lnpost_vectorized(self, paramss) -> np.ndarray:
        a = np.asarray(paramss, dtype=np.float64)
        if self.device_priors:          # priors evaluated on the GPU inside the same call
            return self._batch(self._lib.vela_b200_lnpost_batch, a, None)
        lnpr = np.ascontiguousarray(_priors.lnprior(self.model.priors, a), dtype=np.float64)
        return self._batch(self._lib.vela_b200_lnpost_batch, a, _dp(lnpr))

    def lnpost(self, params) -> float:
        return float(self.lnpost_vectorized(self._free(params)[None, :])[0])

    # -- residuals
    def resids_and_Ninvdiag(self, params):
        p = self._free(params)
        n = int(self._lib.vela_b200_n_rows(self._h))
        y = np.empty(n)
        w = np.empty(n)
        _lib.check(self._lib.vela_b200_residuals(self._h, _dp(p), len(p), _dp(y), _dp(w)), "vela_b200_residuals")
        return y, w

    def phases(self, params):
        """(phase_hi, phase_lo, doppler-shifted spin frequency) per TOA; phase = phase_residual - tzrphase."""
        p = self._free(params)
        n = len(self.toas)
        hi, lo, f = np.empty(n), np.empty(n), np.empty(n)
        _lib.check(self._lib.vela_b200_phases(self._h, _dp(p), len(p), _dp(hi), _dp(lo), _dp(f)), "vela_b200_phases")
        return hi, lo, f

    def noise_weights_inv(self, params) -> np.ndarray:
        p = self._free(params)
        out = np.empty(int(self._lib.vela_b200_n_basis(self._h)))
        _lib.check(self._lib.vela_b200_noise_weights_inv(self._h, _dp(p), len(p), _dp(out)),
                   "vela_b200_noise_weights_inv")
        return out

    def sigma_and_u(self, params):
        """(Sigma^-1 (p, p) symmetric, u (p,)) — _calc_Σinv__and__MT_Ninv_y, gls_likelihood.jl:101-169."""
        p = self._free(params)
        n = int(self._lib.vela_b200_n_basis(self._h))
        S = np.empty((n, n))
        u = np.empty(n)
        _lib.check(self._lib.vela_b200_sigma_and_u(self._h, _dp(p), len(p), _dp(S), _dp(u)), "vela_b200_sigma_and_u")
        return S, u

    # -- conditional distribution of the marginalised amplitudes (pyvela/pyvela/spnta.py:376-465)
    def marginalized_offset_mean_vectorized(self, paramss, with_cf: bool = False, with_lnlike: bool = False):
        """ahat (B, p) [, upper factor R of Sigma^-1 (B, p, p)] [, ln L (B,)] for a batch, from the same
        device factorisation the likelihood uses."""
        a = np.ascontiguousarray(paramss, dtype=np.float64)
        if a.ndim != 2:
            raise ValueError("paramss must be (nwalkers, ndim)")
        B, nd = a.shape
        p = int(self._lib.vela_b200_n_basis(self._h))
        ahat = np.empty((B, p))
        cf = np.empty((B, p, p)) if with_cf else None
        lnl = np.empty(B) if with_lnlike else None
        _lib.check(self._lib.vela_b200_marginalized_mean_batch(
            self._h, _dp(a), B, nd, nd, 1, _dp(ahat), _dp(cf) if with_cf else None,
            _dp(lnl) if with_lnlike else None), "vela_b200_marginalized_mean_batch")
        out = (ahat,) + ((cf,) if with_cf else ()) + ((lnl,) if with_lnlike else ())
        return out[0] if len(out) == 1 else out

    @property
    def has_marginalized_gp_noise(self) -> bool:
        return int(self._lib.vela_b200_n_basis(self._h)) > 0

    def get_marginalized_param_offset_mean_and_covinvcf(self, params):
        """spnta.py:376-394 — (ahat, Sigmainv_cf) for one sample; empty arrays without marginalised GPs."""
        if not self.has_marginalized_gp_noise:
            return np.array([]), np.array([[]])
        ahat, cf = self.marginalized_offset_mean_vectorized(self._free(params)[None, :], with_cf=True)
        return ahat[0], cf[0]

    def get_marginalized_param_offset_mean(self, params):
        return self.get_marginalized_param_offset_mean_and_covinvcf(params)[0]

    def get_marginalized_param_std(self, params):
        """spnta.py:407-417 — sqrt(diag(Sigma)) from the triangular inverse of the factor."""
        if not self.has_marginalized_gp_noise:
            return np.array([])
        from scipy.linalg import solve_triangular
        cf = self.get_marginalized_param_offset_mean_and_covinvcf(params)[1]
        inv = solve_triangular(cf, np.eye(cf.shape[0]), lower=False)
        return np.sqrt(np.sum(inv * inv, axis=1))

    def get_marginalized_param_offset_sample(self, params, rng=None):
        """spnta.py:419-437 — one draw a ~ N(ahat, Sigma) and its log density."""
        if not self.has_marginalized_gp_noise:
            return np.array([]), 0.0
        from scipy.linalg import solve_triangular
        ahat, cf = self.get_marginalized_param_offset_mean_and_covinvcf(params)
        z = (rng if rng is not None else np.random.default_rng()).standard_normal(len(ahat))
        a = ahat + solve_triangular(cf, z, lower=False)
        lnp = -0.5 * (z @ z) + np.sum(np.log(np.diag(cf))) - 0.5 * len(ahat) * np.log(2 * np.pi)
        return a, float(lnp)

    def get_marginalized_gp_noise_realization(self, params, rng=None):
        """spnta.py:456-465 — M a for one draw of the marginalised amplitudes (zeros without GPs)."""
        if not self.has_marginalized_gp_noise:
            return np.zeros(len(self.toas))
        return np.asarray(self.model.kernel.noise_basis) @ self.get_marginalized_param_offset_sample(params, rng)[0]

    # -- residual post-processing (pyvela/pyvela/spnta.py:506-556); y, N^-1 come from the device chain
    def time_residuals(self, params) -> np.ndarray:
        return self.resids_and_Ninvdiag(params)[0][: len(self.toas)]

    def dm_residuals(self, params) -> np.ndarray:
        if not self.toas.wideband:
            raise AssertionError("This method is only defined for wideband datasets.")
        return self.resids_and_Ninvdiag(params)[0][len(self.toas):]

    def scaled_toa_uncertainties(self, params) -> np.ndarray:
        return 1.0 / np.sqrt(self.resids_and_Ninvdiag(params)[1][: len(self.toas)])

    def scaled_dm_uncertainties(self, params) -> np.ndarray:
        """spnta.py:605-620 (wideband only)."""
        if not self.toas.wideband:
            raise AssertionError("This method is only defined for wideband datasets.")
        return 1.0 / np.sqrt(self.resids_and_Ninvdiag(params)[1][len(self.toas):])

    def model_dm(self, params) -> np.ndarray:
        """spnta.py:622-645, in dmu.  Wideband: the model DM the chain accumulates (measured DM minus DM residual);
        the narrowband variant re-runs the dispersion components on the host in the reference and is not mirrored."""
        if not self.toas.wideband:
            raise NotImplementedError("model_dm is mirrored for wideband TOAs only")
        return (self.toas.dm - self.dm_residuals(params)) * 2.41e-16

    def maxpost_params(self, x0=None) -> np.ndarray:
        """spnta.py:647-655: Nelder-Mead on -lnpost from the default parameters."""
        from scipy.optimize import minimize
        start = self.model.read_param_values_to_vector() if x0 is None else np.asarray(x0, dtype=np.float64)
        return minimize(lambda x: -self.lnpost(x), start, method="Nelder-Mead").x

    def whitened_time_residuals(self, params, rng=None) -> np.ndarray:
        """spnta.py:523-556: subtract one GP realisation, then the conditional mean of the ECORR jitter.  The
        ECORR design matrix is block-orthogonal, so its normal equations decouple into one scalar per group."""
        y, w = self.resids_and_Ninvdiag(params)
        n = len(self.toas)
        r = y[:n].copy()
        if self.has_marginalized_gp_noise:
            r -= self.get_marginalized_gp_noise_realization(params, rng)[:n]
        kernel = self.model.kernel
        inner = getattr(kernel, "inner_kernel", kernel)
        groups = getattr(inner, "ecorr_groups", None)
        if groups is not None:
            ph = self.model.param_handler
            full = np.array(ph._default_values, dtype=np.float64)
            full[np.asarray(ph._free_indices0)] = self._free(params)
            ecorr0 = ph.offset("ECORR") if ph.has("ECORR") else 0
            for a0, a1, idx in np.asarray(groups, dtype=np.int64).reshape(-1, 3):
                if idx > 0:
                    sl = slice(a0 - 1, a1)
                    c2 = full[ecorr0 + idx - 1] ** 2
                    r[sl] -= np.sum(r[sl] * w[sl]) / (1.0 / c2 + np.sum(w[sl]))
        return r

    def whitened_dm_residuals(self, params, rng=None) -> np.ndarray:
        d = self.dm_residuals(params)
        if self.has_marginalized_gp_noise:
            d = d - self.get_marginalized_gp_noise_realization(params, rng)[len(self.toas):]
        return d

    # -- benchmarking hooks
    def lnlike_device(self, d_params_ptr: int, B: int, d_out_ptr: int, stream: int = 0):
        _lib.check(self._lib.vela_b200_lnlike_batch_device(self._h, C.c_void_p(d_params_ptr), B,
                                                           C.c_void_p(d_out_ptr), C.c_void_p(stream)),
                   "vela_b200_lnlike_batch_device")

    def lnpost_device(self, d_params_ptr: int, B: int, d_lnprior_ptr: int, d_out_ptr: int, stream: int = 0):
        _lib.check(self._lib.vela_b200_lnpost_batch_device(self._h, C.c_void_p(d_params_ptr), B,
                                                           C.c_void_p(d_lnprior_ptr), 1, C.c_void_p(d_out_ptr),
                                                           C.c_void_p(stream)), "vela_b200_lnpost_batch_device")

    def chain_specialized(self, on=None) -> bool:
        """Whether the per-model (NVRTC) chain kernel is in use; `on=True/False` switches kernels (A/B tests)."""
        return bool(self._lib.vela_b200_chain_specialized(self._h, -1 if on is None else int(bool(on))))

    def sigma_structured(self) -> bool:
        """True when Sigma^-1 is assembled from trigonometric column sums (harmonic Fourier basis detected)."""
        return bool(self._lib.vela_b200_sigma_structured(self._h))

    def sigma_info(self, B: int) -> dict:
        """What the Sigma stage executes for a batch of B walkers (measurement bookkeeping): path taken and executed flops."""
        import ctypes as C
        info = (C.c_int64 * 8)()
        fl = float(self._lib.vela_b200_sigma_info(self._h, int(B), info))
        keys = ("structured", "two_stage", "table_columns", "rows_padded", "stage1_columns", "stage2_columns", "moments_per_block", "cells")
        d = dict(zip(keys, (int(v) for v in info)))
        d["executed_flops_per_walker"] = fl
        return d

    def jit_log(self) -> str:
        return (self._lib.vela_b200_jit_log(self._h) or b"").decode(errors="replace")

    def launch_count(self) -> int:
        return int(self._lib.vela_b200_launch_count(self._h))

    def set_stage_timing(self, on: bool):
        self._lib.vela_b200_set_stage_timing(self._h, int(on))

    def last_stage_ms(self):
        return [float(self._lib.vela_b200_last_stage_ms(self._h, s)) for s in range(5)]


_CACHE: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


def _pulsar(*args):
    """(pulsar, params) | (model, toas, params) -> (Pulsar, params)"""
    if isinstance(args[0], Pulsar):
        return args[0], args[1]
    model, toas, params = args
    per_model = _CACHE.setdefault(model, {})
    psr = per_model.get(id(toas))
    if psr is None or psr.toas is not toas:
        psr = Pulsar(model, toas)
        per_model[id(toas)] = psr
    return psr, params


def calc_lnlike(*args) -> float:
    psr, params = _pulsar(*args)
    return psr.lnlike(params)


calc_lnlike_serial = calc_lnlike


def calc_lnlike_vectorized(*args) -> np.ndarray:
    psr, paramss = _pulsar(*args)
    return psr.lnlike_vectorized(paramss)


def calc_lnprior(obj, params):
    model = obj.model if isinstance(obj, Pulsar) else obj
    p = np.asarray(params, dtype=np.float64)
    ph = model.param_handler
    if p.ndim == 1 and p.shape[0] == len(ph._default_values) and p.shape[0] != ph._nfree:
        p = ph.read_param_values_to_vector(p)
    return _priors.lnprior(model.priors, p)


def prior_transform(obj, cube):
    model = obj.model if isinstance(obj, Pulsar) else obj
    return _priors.prior_transform(model.priors, cube)


def calc_lnpost(*args) -> float:
    psr, params = _pulsar(*args)
    return psr.lnpost(params)


calc_lnpost_serial = calc_lnpost


def calc_lnpost_vectorized(*args) -> np.ndarray:
    psr, paramss = _pulsar(*args)
    return psr.lnpost_vectorized(paramss)


def calc_chi2(*args) -> float:
    psr, params = _pulsar(*args)
    return psr.chi2(params)


calc_chi2_serial = calc_chi2


def degrees_of_freedom(model: TimingModel, toas: TOAs) -> int:
    """wls_chi2.jl:85-88"""
    return (2 if toas.wideband else 1) * len(toas) - model.param_handler._nfree


def calc_chi2_reduced(model, toas, params) -> float:
    return calc_chi2(model, toas, params) / degrees_of_freedom(model, toas)


def form_residuals(*args):
    """residuals.jl:19-26 (narrowband: time residuals) / wideband_residuals.jl:24-27 ((time, DM) pairs)."""
    psr, params = _pulsar(*args)
    y, _ = psr.resids_and_Ninvdiag(params)
    n = len(psr.toas)
    return (y[:n], y[n:]) if psr.toas.wideband else y


def _calc_resids_and_Ninvdiag(*args):
    psr, params = _pulsar(*args)
    return psr.resids_and_Ninvdiag(params)


def calc_noise_weights_inv(*args):
    psr, params = _pulsar(*args)
    return psr.noise_weights_inv(params)


def get_lnlike_serial_func(model, toas) -> Callable:
    psr, _ = _pulsar(model, toas, None)
    return psr.lnlike


get_lnlike_parallel_func = get_lnlike_serial_func
get_lnlike_func = get_lnlike_serial_func


def get_chi2_serial_func(model, toas) -> Callable:
    psr, _ = _pulsar(model, toas, None)
    return psr.chi2


get_chi2_parallel_func = get_chi2_serial_func
get_chi2_func = get_chi2_serial_func


def get_lnprior_func(model) -> Callable:
    return lambda params: calc_lnprior(model, params)


def get_prior_transform_func(model) -> Callable:
    return lambda cube: prior_transform(model, cube)


def get_lnpost_func(model, toas, vectorize: bool = False) -> Callable:
    """posterior.jl:34-46"""
    psr, _ = _pulsar(model, toas, None)
    return psr.lnpost_vectorized if vectorize else psr.lnpost
