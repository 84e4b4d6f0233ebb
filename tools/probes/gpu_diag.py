"""Diagnostic (not a test): GPU-vs-oracle differences on the golden fixtures."""
import sys, os
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle'); sys.path.insert(0, 'tests')
from __graft_entry__ import load_package
vb = load_package()
import oracle as O
from conftest import Golden, GOLDEN_NAMES
print('lib', vb.lib.LIB_PATH)
for name in GOLDEN_NAMES:
    g = Golden(vb, name)
    psr = vb.Pulsar(g.model, g.toas)
    y, w = psr.resids_and_Ninvdiag(g.x0)
    yo, wo = O.residuals(g.md, g.x0)
    n = len(g.toas)
    got = psr.lnlike_vectorized(g.walkers); want = O.lnlike_batch(g.md, g.walkers)
    rel = np.abs(got - want) / np.abs(want)
    print(f"{name:12s} max|dy|={np.max(np.abs(y[:n]-yo[:n])):.3e}s  max rel dw={np.max(np.abs(w/wo-1)):.2e}  "
          f"lnlike: max abs diff={np.max(np.abs(got-want)):.3e} max rel={rel.max():.3e} median rel={np.median(rel):.2e} "
          f"at default: {got[0]!r} vs {want[0]!r}")
    if g.md.n_basis:
        S, u = psr.sigma_and_u(g.x0)
        So, uo = O.sigma(g.md, g.x0)
        print('   Sigma rel err', np.max(np.abs(S - So) / np.abs(So).max()), ' u rel err', np.max(np.abs(u - uo)) / np.abs(uo).max(),
              'diag rel', np.max(np.abs(np.diag(S) / np.diag(So) - 1)))
        bad = np.argwhere(~np.isfinite(S))
        print('   non-finite entries', len(bad), bad[:5].tolist(), 'u nonfinite', np.argwhere(~np.isfinite(u))[:5].ravel().tolist())
import ctypes as C, mpmath
mpmath.mp.prec = 200
rng = np.random.default_rng(1)
x = np.concatenate([rng.uniform(-7, 7, 2000), rng.uniform(-1e3, 1e3, 200), [0.0, 1e-300, np.pi/2, np.pi, 4.663868852922854]])
s = np.empty_like(x); c = np.empty_like(x)
dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
vb.lib.lib().vela_b200_debug_sincos(dp(x), len(x), dp(s), dp(c))
bad = 0
for xi, si, ci in zip(x, s, c):
    if float(mpmath.sin(mpmath.mpf(xi))) != si or float(mpmath.cos(mpmath.mpf(xi))) != ci: bad += 1
print('sincos_cr: not correctly rounded in', bad, 'of', len(x), '; glibc disagrees with mpmath in',
      sum(1 for xi in x if float(mpmath.sin(mpmath.mpf(xi))) != np.sin(xi) or float(mpmath.cos(mpmath.mpf(xi))) != np.cos(xi)))
