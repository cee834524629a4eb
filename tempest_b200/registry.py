"""Registry priors and likelihoods that the fused CUDA mutation kernel evaluates in-kernel.

The reference takes arbitrary Python callables (``prior_transform(u_row) -> x_row`` per row,
tempest/steps/mutate.py:103, tempest/mcmc.py:157; ``log_likelihood(x[N,D]) -> logl[N]``
batched, tempest/core.py:321-322).  A registry object *is* such a callable -- its
``__call__`` is an ordinary numpy function, so the very same object can be handed to the
reference or to the CPU oracle -- and it also carries ``(kernel_id, dparams)`` which
``csrc/tb_like.cuh`` evaluates with the same operation order (explicit ``__dmul_rn`` /
``__dadd_rn``, no FMA contraction), so log-likelihoods agree bit-for-bit given equal ``x``
wherever only + - * / are involved.

The numpy forms are written as explicit left-to-right accumulations so the order of
floating-point operations is unambiguous (``np.sum`` switches to 8-lane pairwise blocks
for >= 8 terms).  For the README example (D = 10) ``Rosenbrock`` is bit-identical to the
README's ``-np.sum(10*(x[:,::2]**2 - x[:,1::2])**2 + (x[:,::2]-1)**2, axis=1)``.
"""

from __future__ import annotations

import math
from typing import Sequence

import numpy as np

# kernel ids -- keep in sync with include/tempest_b200.h (TB_LIKE_*, TB_PRIOR_*)
LIKE_ROSENBROCK = 0
LIKE_GAUSSIAN = 1
LIKE_ISO_MIXTURE = 2
LIKE_TWIN_SHELLS = 3
PRIOR_AFFINE = 0


def _np_logaddexp_scalar_form(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    return np.logaddexp(a, b)


class UniformPrior:
    """``x = lo + (hi - lo) * u`` elementwise (so per-row and batched calls agree)."""

    kernel_id = PRIOR_AFFINE

    def __init__(self, lo, hi, n_dim: int):
        self.n_dim = int(n_dim)
        self.lo = np.broadcast_to(np.asarray(lo, dtype=float), (self.n_dim,)).copy()
        self.hi = np.broadcast_to(np.asarray(hi, dtype=float), (self.n_dim,)).copy()
        self.scale = self.hi - self.lo

    def __call__(self, u: np.ndarray) -> np.ndarray:
        return self.lo + self.scale * np.asarray(u, dtype=float)

    def dparams(self) -> np.ndarray:
        """[lo[D], scale[D]]"""
        return np.concatenate([self.lo, self.scale])

    def log_volume(self) -> float:
        return float(np.sum(np.log(self.scale)))


class _Likelihood:
    kernel_id = -1
    n_dim = 0

    def dparams(self) -> np.ndarray:  # pragma: no cover - interface
        raise NotImplementedError

    def __call__(self, x):  # pragma: no cover - interface
        raise NotImplementedError


class Rosenbrock(_Likelihood):
    """``logL = -sum_i [a (x_{2i}^2 - x_{2i+1})^2 + (x_{2i} - 1)^2]`` (README.md:44-71, a=10)."""

    kernel_id = LIKE_ROSENBROCK

    def __init__(self, n_dim: int, a: float = 10.0):
        if n_dim % 2:
            raise ValueError("Rosenbrock needs an even n_dim")
        self.n_dim = int(n_dim)
        self.a = float(a)

    def __call__(self, x: np.ndarray) -> np.ndarray:
        x = np.atleast_2d(np.asarray(x, dtype=float))
        acc = None
        for i in range(self.n_dim // 2):
            xe = x[:, 2 * i]
            xo = x[:, 2 * i + 1]
            t = xe * xe - xo
            t = self.a * (t * t)
            s = xe - 1.0
            term = t + s * s
            acc = term if acc is None else acc + term
        return -acc

    def dparams(self) -> np.ndarray:
        return np.array([self.a])


class GaussianLikelihood(_Likelihood):
    """``N(x; mean, cov)`` evaluated as ``y = Linv (x - mean)``, ``-0.5 sum y^2 + const`` with
    ``Linv`` the inverse lower Cholesky factor (row-by-row, left-to-right dot products)."""

    kernel_id = LIKE_GAUSSIAN

    def __init__(self, mean: Sequence[float], cov: np.ndarray):
        self.mean = np.asarray(mean, dtype=float).copy()
        self.n_dim = self.mean.size
        cov = np.asarray(cov, dtype=float)
        if cov.ndim == 1:
            cov = np.diag(cov)
        chol = np.linalg.cholesky(cov)
        self.linv = np.linalg.inv(chol)
        self.linv = np.tril(self.linv)
        self.const = float(-np.sum(np.log(np.diag(chol))) - 0.5 * self.n_dim * math.log(2.0 * math.pi))

    @classmethod
    def ar1(cls, n_dim: int, rho: float = 0.5):
        """zero mean, ``cov_ij = rho^|i-j|``."""
        i = np.arange(n_dim)
        return cls(np.zeros(n_dim), rho ** np.abs(i[:, None] - i[None, :]))

    def __call__(self, x: np.ndarray) -> np.ndarray:
        x = np.atleast_2d(np.asarray(x, dtype=float))
        diff = x - self.mean
        acc = None
        for i in range(self.n_dim):
            y = self.linv[i, 0] * diff[:, 0]
            for j in range(1, i + 1):
                y = y + self.linv[i, j] * diff[:, j]
            sq = y * y
            acc = sq if acc is None else acc + sq
        return -0.5 * acc + self.const

    def dparams(self) -> np.ndarray:
        """[const, mean[D], Linv[D*D] row-major]"""
        return np.concatenate([[self.const], self.mean, self.linv.ravel()])


class IsotropicMixture(_Likelihood):
    """``logL = log sum_k w_k N(x; mu_k, v_k I)`` as ``m + log(sum_k exp(lp_k - m))`` with
    ``lp_k = c_k - r_k^2 * h_k``, ``c_k = log w_k - D/2 log(2 pi v_k)``, ``h_k = 1/(2 v_k)``."""

    kernel_id = LIKE_ISO_MIXTURE

    def __init__(self, means: np.ndarray, variances: Sequence[float], weights: Sequence[float]):
        self.means = np.atleast_2d(np.asarray(means, dtype=float)).copy()
        self.K, self.n_dim = self.means.shape
        v = np.broadcast_to(np.asarray(variances, dtype=float), (self.K,))
        w = np.broadcast_to(np.asarray(weights, dtype=float), (self.K,))
        self.c = np.log(w) - 0.5 * self.n_dim * np.log(2.0 * np.pi * v)
        self.h = 1.0 / (2.0 * v)

    @classmethod
    def four_corners(cls, n_dim: int = 2, sep: float = 4.0, var: float = 0.25):
        """means (+-sep, +-sep), variance ``var``, weights 1/4."""
        means = np.zeros((4, n_dim))
        means[:, :2] = [[sep, sep], [sep, -sep], [-sep, sep], [-sep, -sep]]
        return cls(means, var, 0.25)

    def __call__(self, x: np.ndarray) -> np.ndarray:
        x = np.atleast_2d(np.asarray(x, dtype=float))
        lps = []
        for k in range(self.K):
            r2 = None
            for j in range(self.n_dim):
                dlt = x[:, j] - self.means[k, j]
                sq = dlt * dlt
                r2 = sq if r2 is None else r2 + sq
            lps.append(self.c[k] - r2 * self.h[k])
        m = lps[0]
        for k in range(1, self.K):
            m = np.maximum(m, lps[k])
        s = np.exp(lps[0] - m)
        for k in range(1, self.K):
            s = s + np.exp(lps[k] - m)
        return m + np.log(s)

    def dparams(self) -> np.ndarray:
        """[K, c[K], h[K], means[K*D]]"""
        return np.concatenate([[float(self.K)], self.c, self.h, self.means.ravel()])


class TwinShells(_Likelihood):
    """``logL = logaddexp(S(x;c1), S(x;c2))``, ``S = -(|x-c| - r)^2/(2 w^2) - 0.5 log(2 pi w^2)``
    (SURVEY App. D config C5)."""

    kernel_id = LIKE_TWIN_SHELLS

    def __init__(self, n_dim: int, offset: float = 3.5, radius: float = 2.0, width: float = 0.1):
        self.n_dim = int(n_dim)
        self.c1 = np.zeros(n_dim)
        self.c2 = np.zeros(n_dim)
        self.c1[0] = -offset
        self.c2[0] = offset
        self.r = float(radius)
        self.h = 1.0 / (2.0 * width * width)
        self.const = -0.5 * math.log(2.0 * math.pi * width * width)

    def _shell(self, x, c):
        r2 = None
        for j in range(self.n_dim):
            dlt = x[:, j] - c[j]
            sq = dlt * dlt
            r2 = sq if r2 is None else r2 + sq
        t = np.sqrt(r2) - self.r
        return self.const - (t * t) * self.h

    def __call__(self, x: np.ndarray) -> np.ndarray:
        x = np.atleast_2d(np.asarray(x, dtype=float))
        return np.logaddexp(self._shell(x, self.c1), self._shell(x, self.c2))

    def dparams(self) -> np.ndarray:
        """[r, h, const, c1[D], c2[D]]"""
        return np.concatenate([[self.r, self.h, self.const], self.c1, self.c2])


def is_registry_likelihood(obj) -> bool:
    return isinstance(obj, _Likelihood)


def is_registry_prior(obj) -> bool:
    return isinstance(obj, UniformPrior)
