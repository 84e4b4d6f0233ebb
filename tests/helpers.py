"""Test helpers: an oracle-backed evaluator for the synthetic-data generator (tests only)."""
import numpy as np


class OracleEval:
    def __init__(self, vb, O, model, toas):
        self.O = O
        self.md = vb.ModelDescriptor(model, toas)

    def phases(self, x):
        hi, lo, f, _, _ = self.O.phases(self.md, x)
        return hi, lo, f

    def resids_and_Ninvdiag(self, x):
        return self.O.residuals(self.md, x)


def oracle_eval_factory(vb, O):
    return lambda model, toas: OracleEval(vb, O, model, toas)


# ---- host build of the product's device arithmetic (tests/host_emul): the chain kernel's per-(TOA, walker) math on the CPU
_EMUL = None


def emul_lib():
    import ctypes as C
    import os
    import subprocess
    global _EMUL
    if _EMUL is None:
        here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emul")
        subprocess.run(["make", "-C", here], check=True, stdout=subprocess.DEVNULL)
        L = C.CDLL(os.path.join(here, "libvela_emul.so"))
        dp = C.POINTER(C.c_double)
        L.emul_chain.argtypes = [C.c_void_p, dp, C.c_long, C.c_int, C.c_int, dp, dp, dp, dp, dp]
        L.emul_spec_check.argtypes = [C.c_void_p, dp, C.c_long, C.c_int, C.POINTER(C.c_long)]
        _EMUL = L
    return _EMUL


def emul_chain(md, paramss, emit=1, fast=False):
    """Runs prologue + chain + residual terms of the CUDA source on the CPU.
    fast: the fast-arithmetic build of the chain (the product's default) instead of the strict one.
    emit 0/1: returns (y[B, n_rows], w[B, n_rows], chi2[B], logdetN[B]); emit 2: (phase_hi, phase_lo, f, chi2, logdetN)."""
    import ctypes as C
    p = np.ascontiguousarray(np.atleast_2d(paramss), dtype=np.float64)
    B = p.shape[0]
    dp = C.POINTER(C.c_double)
    Y = np.zeros((B, md.n_rows)); W = np.zeros((B, md.n_rows)); F = np.zeros((B, md.n_rows))
    chi = np.zeros(B); lg = np.zeros(B)
    ptr = lambda a: a.ctypes.data_as(dp)  # noqa: E731
    rc = emul_lib().emul_chain(C.addressof(md.desc), ptr(p), B, emit, int(bool(fast)), ptr(Y), ptr(W), ptr(F), ptr(chi), ptr(lg))
    assert rc == 0
    if emit == 2:
        return Y, W, F, chi, lg
    return Y, W, chi, lg


def emul_spec_check(md, paramss, emit=1):
    """Speculative vs branching evaluation of the fast chain build on the CPU (tests/host_emul/emul.cpp):
    returns (pairs, pairs that raised `redo`, pairs without `redo` whose outputs differ in any bit)."""
    import ctypes as C
    p = np.ascontiguousarray(np.atleast_2d(paramss), dtype=np.float64)
    counts = (C.c_long * 3)()
    rc = emul_lib().emul_spec_check(C.addressof(md.desc), p.ctypes.data_as(C.POINTER(C.c_double)), p.shape[0], emit, counts)
    assert rc == 0
    return tuple(counts)
