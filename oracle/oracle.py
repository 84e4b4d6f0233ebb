"""ctypes binding of the CPU oracle (oracle/vela_oracle.c).  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package (vela.jl_b200/) never does.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libvela_oracle.so")
    src = os.path.join(_HERE, "vela_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "vela_b200.h")
    stale = (not os.path.exists(so)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(so) for p in (src, hdr))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libvela_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return so


def build_native() -> str:
    """The same source built `-O3 -march=native` (SURVEY 8d's CPU-baseline flags) ON the machine that runs it; contraction
    stays off, so the results are the standard build's bit for bit.  Used by bench.py's CPU legs only."""
    out_dir = os.path.join(_HERE, "_native")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libvela_oracle_native.so")
    src = os.path.join(_HERE, "vela_oracle.c")
    cc = "/usr/bin/gcc" if os.access("/usr/bin/gcc", os.X_OK) else "gcc"
    subprocess.run([cc, "-O3", "-march=native", "-fPIC", "-fopenmp", "-ffp-contract=off", "-fno-fast-math",
                    "-fvisibility=hidden", "-shared", "-o", so, src, "-lm"], check=True, stdout=subprocess.DEVNULL,
                   stderr=subprocess.DEVNULL)
    return so


_NATIVE = False


def use_native() -> bool:
    """Switch this module to the -O3 -march=native build (compiled now, on this host).  Returns False (and keeps the
    standard build) when the compile fails."""
    global _LIB, _NATIVE
    try:
        so = build_native()
    except Exception:
        return False
    _LIB = None
    _NATIVE = True
    lib(so)
    return True


def lib(path: str = None):
    global _LIB
    if _LIB is None:
        L = C.CDLL(path or build())
        dp = C.POINTER(C.c_double)
        L.oracle_lnlike_batch.argtypes = [C.c_void_p, dp, C.c_int64, C.c_int64, C.c_int64, dp, C.c_int, C.c_int]
        L.oracle_lnlike_batch.restype = C.c_int
        L.oracle_residuals.argtypes = [C.c_void_p, dp, dp, dp]
        L.oracle_phases.argtypes = [C.c_void_p, dp, dp, dp, dp, dp, dp]
        L.oracle_noise_weights_inv.argtypes = [C.c_void_p, dp, dp]
        L.oracle_sigma.argtypes = [C.c_void_p, dp, dp, dp]
        L.oracle_gls_lnlike_raw.argtypes = [dp, C.c_int64, C.c_int64, dp, dp, dp, C.POINTER(C.c_uint64), C.c_int, dp]
        L.oracle_gls_lnlike_raw.restype = C.c_double
        L.oracle_sincos_cr.argtypes = [C.c_double, dp, dp]
        L.oracle_mikkola.argtypes = [C.c_double, C.c_double]
        L.oracle_mikkola.restype = C.c_double
        L.oracle_powerlaw.argtypes = [C.c_double] * 4
        L.oracle_powerlaw.restype = C.c_double
        L.oracle_max_threads.restype = C.c_int
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _vec(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def lnlike_batch(md, paramss, nthreads: int = 0, chi2_only: bool = False) -> np.ndarray:
    """calc_lnlike_vectorized / calc_chi2 over rows of `paramss` (B, nfree)."""
    p = np.ascontiguousarray(np.atleast_2d(paramss), dtype=np.float64)
    assert p.shape[1] == md.desc.n_free
    out = np.empty(p.shape[0])
    rc = lib().oracle_lnlike_batch(C.addressof(md.desc), _dp(p), p.shape[0], p.shape[1], 1, _dp(out),
                                   int(nthreads), int(chi2_only))
    if rc != 0:
        raise RuntimeError(f"oracle_lnlike_batch failed: {rc}")
    return out


def lnlike(md, params) -> float:
    return float(lnlike_batch(md, np.asarray(params)[None, :])[0])


def chi2(md, params) -> float:
    return float(lnlike_batch(md, np.asarray(params)[None, :], chi2_only=True)[0])


def residuals(md, params):
    """_calc_resids_and_Ninvdiag: (y, Ninvdiag), each n_rows long."""
    p = _vec(params)
    y = np.empty(md.n_rows)
    w = np.empty(md.n_rows)
    lib().oracle_residuals(C.addressof(md.desc), _dp(p), _dp(y), _dp(w))
    return y, w


def phases(md, params):
    """(phase_hi, phase_lo, doppler-shifted spin frequency, total delay, tzr phase (hi, lo))."""
    p = _vec(params)
    n = len(md.toas)
    hi, lo, f, dly = (np.empty(n) for _ in range(4))
    tzr = np.empty(2)
    lib().oracle_phases(C.addressof(md.desc), _dp(p), _dp(hi), _dp(lo), _dp(f), _dp(dly), _dp(tzr))
    return hi, lo, f, dly, tzr


def noise_weights_inv(md, params) -> np.ndarray:
    p = _vec(params)
    out = np.empty(md.n_basis)
    lib().oracle_noise_weights_inv(C.addressof(md.desc), _dp(p), _dp(out))
    return out


def sigma(md, params):
    """(Sigma^-1 as a full symmetric (p, p) array, u = M^T N^-1 y)."""
    p = _vec(params)
    n = md.n_basis
    S = np.empty((n, n))
    u = np.empty(n)
    lib().oracle_sigma(C.addressof(md.desc), _dp(p), _dp(S), _dp(u))
    # C side: S[c*p + q] valid for q <= c  ->  numpy S[c, q]; symmetrise
    L = np.tril(S)
    return L + np.tril(S, -1).T, u


def gls_lnlike_raw(Mmat, Ninvdiag, Phiinv, y, ecorr_groups=None, ecorr=None) -> float:
    """_gls_lnlike_serial for a WhiteNoiseKernel, or an EcorrKernel given as (start, stop, index) rows + ECORR values."""
    Mf = np.asfortranarray(Mmat, dtype=np.float64)
    n, p = Mf.shape
    if ecorr_groups is None:
        grp, ng, ec = None, 0, None
    else:
        g = np.ascontiguousarray(ecorr_groups, dtype=np.uint64).reshape(-1, 3)
        grp, ng = g.ctypes.data_as(C.POINTER(C.c_uint64)), len(g)
        e = _vec(ecorr)
        ec = _dp(e)
    return float(lib().oracle_gls_lnlike_raw(_dp(Mf), n, p, _dp(_vec(Ninvdiag)), _dp(_vec(Phiinv)), _dp(_vec(y)), grp, ng, ec))


def sincos_cr(x: float):
    s, c = C.c_double(), C.c_double()
    lib().oracle_sincos_cr(x, C.byref(s), C.byref(c))
    return s.value, c.value


def mikkola(l: float, e: float) -> float:
    return float(lib().oracle_mikkola(l, e))


def powerlaw(A, gamma, f, f1) -> float:
    return float(lib().oracle_powerlaw(A, gamma, f, f1))


def max_threads() -> int:
    return int(lib().oracle_max_threads())
