"""Host-side priors: `lnprior` and `prior_transform` for the default prior families.

The reference evaluates priors per sample on the host with Distributions.jl
(src/prior/prior.jl:13-37, src/prior/simple_prior.jl:7-55, src/prior/known_priors.jl:8-100) and adds
the result to the likelihood (src/likelihood/posterior.jl:9-12).  In the drop-in the Julia host keeps
doing exactly that; this numpy mirror plays the same role for the Python harness, vectorised over
the walker batch.  It is O(B * ndim) bookkeeping, not part of the GPU hot path.
"""

from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Any, Sequence

import numpy as np
from scipy import special

DEFAULT_PRIOR, CHEAT_PRIOR, USER_DEFINED_PRIOR = 0, 1, 2    # simple_prior.jl:3


class Distribution:
    def logpdf(self, x):
        raise NotImplementedError

    def quantile(self, q):
        raise NotImplementedError


@dataclass
class Uniform(Distribution):
    a: float
    b: float

    def logpdf(self, x):
        x = np.asarray(x, dtype=float)
        return np.where((x >= self.a) & (x <= self.b), -math.log(self.b - self.a), -np.inf)

    def quantile(self, q):
        return self.a + np.asarray(q) * (self.b - self.a)


@dataclass
class Normal(Distribution):
    mu: float
    sigma: float

    def logpdf(self, x):
        z = (np.asarray(x, dtype=float) - self.mu) / self.sigma
        return -0.5 * z * z - math.log(self.sigma) - 0.5 * math.log(2 * math.pi)

    def quantile(self, q):
        return self.mu + self.sigma * math.sqrt(2) * special.erfinv(2 * np.asarray(q) - 1)


@dataclass
class LogNormal(Distribution):
    mu: float
    sigma: float

    def logpdf(self, x):
        x = np.asarray(x, dtype=float)
        with np.errstate(divide="ignore", invalid="ignore"):
            lx = np.log(x)
            z = (lx - self.mu) / self.sigma
            out = -0.5 * z * z - math.log(self.sigma) - 0.5 * math.log(2 * math.pi) - lx
        return np.where(x > 0, out, -np.inf)

    def quantile(self, q):
        return np.exp(self.mu + self.sigma * math.sqrt(2) * special.erfinv(2 * np.asarray(q) - 1))


@dataclass
class LogUniform(Distribution):
    a: float
    b: float

    def logpdf(self, x):
        x = np.asarray(x, dtype=float)
        with np.errstate(divide="ignore", invalid="ignore"):
            out = -np.log(x) - math.log(math.log(self.b / self.a))
        return np.where((x >= self.a) & (x <= self.b), out, -np.inf)

    def quantile(self, q):
        return np.exp(math.log(self.a) + np.asarray(q) * math.log(self.b / self.a))


@dataclass
class TruncatedNormal(Distribution):
    """Distributions.truncated(Normal(mu, sigma), lower, upper) — what pyvela builds for bounded custom priors
    (pyvela/pyvela/priors.py:245-256)."""
    mu: float
    sigma: float
    lower: float
    upper: float

    def _mass(self):
        return special.ndtr((self.upper - self.mu) / self.sigma) - special.ndtr((self.lower - self.mu) / self.sigma)

    def logpdf(self, x):
        x = np.asarray(x, dtype=float)
        z = (x - self.mu) / self.sigma
        out = -0.5 * z * z - math.log(self.sigma) - 0.5 * math.log(2 * math.pi) - math.log(self._mass())
        return np.where((x >= self.lower) & (x <= self.upper), out, -np.inf)

    def quantile(self, q):
        c0 = special.ndtr((self.lower - self.mu) / self.sigma)
        return self.mu + self.sigma * special.ndtri(c0 + np.asarray(q) * self._mass())


class _Custom(Distribution):
    lb, ub = 0.0, 1.0

    def pdf_expr(self, x):
        raise NotImplementedError

    def logpdf(self, x):   # known_priors.jl:10-11
        x = np.asarray(x, dtype=float)
        inside = (x >= self.lb) & (x <= self.ub)
        with np.errstate(divide="ignore", invalid="ignore"):
            out = np.log(self.pdf_expr(np.where(inside, x, 0.5 * (self.lb + min(self.ub, self.lb + 1.0)))))
        return np.where(inside, out, -np.inf)


class KINPriorDistribution(_Custom):          # known_priors.jl:33-37
    lb, ub = 0.0, math.pi
    pdf_expr = staticmethod(lambda i: np.sin(i) / 2)
    quantile = staticmethod(lambda q: np.arccos(1 - 2 * np.asarray(q)))


class SINIPriorDistribution(_Custom):         # known_priors.jl:45-49
    pdf_expr = staticmethod(lambda s: s / np.sqrt(1 - s * s))
    quantile = staticmethod(lambda q: np.sqrt((2 - np.asarray(q)) * np.asarray(q)))


class STIGMAPriorDistribution(_Custom):       # known_priors.jl:57-61
    pdf_expr = staticmethod(lambda s: 4 * s / (1 + s**2) ** 2)
    quantile = staticmethod(lambda q: np.asarray(q) / (2 - np.asarray(q)))


class SHAPMAXPriorDistribution(_Custom):      # known_priors.jl:69-73
    lb, ub = 0.0, math.inf
    pdf_expr = staticmethod(lambda S: (1 - np.exp(-S)) / np.sqrt(2 * np.exp(S) - 1))
    quantile = staticmethod(lambda q: np.log((np.sqrt(2 * np.asarray(q) - np.asarray(q) ** 2) + 1)
                                             / (np.asarray(q) - 1) ** 2))


class LatitudePriorDistribution(_Custom):     # known_priors.jl:96-100
    lb, ub = -math.pi / 2, math.pi / 2
    pdf_expr = staticmethod(lambda p: np.cos(p) / 2)
    quantile = staticmethod(lambda q: np.arcsin(2 * np.asarray(q) - 1))


class PXPriorDistribution(_Custom):           # known_priors.jl:82-88
    def __init__(self, Rmax: float):
        self.Rmax = Rmax
        self.lb, self.ub = 1 / Rmax, math.inf

    def pdf_expr(self, x):
        return 3 / (self.Rmax**3 * x**4)

    def quantile(self, q):
        return 1 / (self.Rmax * np.cbrt(1 - np.asarray(q)))


@dataclass
class SimplePrior:
    """simple_prior.jl:17-32 — prior on a single `Parameter`."""
    name: str
    distribution: Distribution
    source_type: int = DEFAULT_PRIOR


@dataclass
class SimplePriorMulti:
    """simple_prior.jl:39-55 — prior on element `index` (1-based) of a `MultiParameter`."""
    name: str
    index: int
    distribution: Distribution
    source_type: int = DEFAULT_PRIOR


def lnprior(priors: Sequence[Any], free_values) -> np.ndarray:
    """prior.jl:13-19 — sum of logpdf over the free parameters; `free_values` is (ndim,) or (B, ndim).
    The priors tuple is ordered like the free parameters (timing_model.jl:40)."""
    x = np.asarray(free_values, dtype=np.float64)
    single = x.ndim == 1
    x = np.atleast_2d(x)
    if x.shape[1] != len(priors):
        raise AssertionError("number of priors must match the number of free parameters")
    total = np.zeros(x.shape[0])
    for k, pr in enumerate(priors):
        total = total + pr.distribution.logpdf(x[:, k])
    return float(total[0]) if single else total


def prior_transform(priors: Sequence[Any], cube) -> np.ndarray:
    """prior.jl:31-32 — quantile of each prior at each cube coordinate; (ndim,) or (B, ndim)."""
    q = np.asarray(cube, dtype=np.float64)
    single = q.ndim == 1
    q = np.atleast_2d(q)
    out = np.empty_like(q)
    for k, pr in enumerate(priors):
        out[:, k] = pr.distribution.quantile(q[:, k])
    return out[0] if single else out
