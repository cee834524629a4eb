"""CPU oracle for the Persistent Sampling hot path -- TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference algorithm (minaskar/tempest v0.2.1,
paths below are relative to /root/reference).  It exists to CHECK the CUDA path:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under ``tempest_b200/`` imports it and
the product path raises when its CUDA library is missing -- there is no CPU fallback.

Parity status: PINNED.  ``oracle/gen_golden.py`` runs the unmodified reference under
``np.random.seed(s)`` and stores its per-iteration outputs in ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` replays this oracle on the same legacy MT19937 stream and
requires the outputs to agree (bit-exact for indices/decisions/betas, <=1e-12 otherwise),
plus the reference's own known-answer vectors (tests/test_mcmc.py:14-94,187-217;
tests/test_tools.py:39-50,88-94; tests/test_steps.py:39-56,100-145;
tests/test_end_to_end.py:20,71-73).

Third-party arithmetic restated here (numpy / scipy are dependencies of the reference, not
vendored in it: pyproject.toml:24,26, uv.lock numpy 2.4.1 / scipy 1.17.0; this image has
numpy 2.3.5 / scipy 1.18.1): legacy ``RandomState.choice`` = cumsum -> /= last ->
searchsorted(right); ``np.cumsum`` strictly sequential; ``np.sum`` pairwise (8-lane blocks
of <=128); ``np.logaddexp.reduce`` sequential; ``np.percentile`` linear interpolation.

Every random variate is drawn through a *stream* object (``LegacyStream`` = numpy's
global MT19937 stream, consumed in exactly the reference's order) which can record the
variates it hands out into per-iteration *tapes*; the CUDA path replays those tapes.
"""

from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

# Algorithm constants (tempest/config.py:233-242)
BETA_TOLERANCE = 1e-4
BETA_RTOL = 1e-8
ESS_TOLERANCE = 0.01
METRIC_ATOL = 0.5
METRIC_ATOL_CV = 0.01
DOF_FALLBACK = 1e6
TRIM_ESS = 0.99
TRIM_BINS = 1000
MAX_BISECTION = 200  # tempest/steps/reweight.py:121
SQRTEPS = math.sqrt(float(np.finfo(np.float64).eps))  # tempest/tools.py:7
_TINY = float(np.finfo(float).tiny)


# --------------------------------------------------------------------------------------
# Random streams and tapes
# --------------------------------------------------------------------------------------
class LegacyStream:
    """numpy's legacy MT19937 stream, i.e. what ``np.random.seed(seed)`` followed by
    ``np.random.rand/randn/random/gamma/choice`` hands to the reference.

    ``np.random.gamma(k, theta)`` is ``theta * standard_gamma(k)`` bit-for-bit, so the
    stream exposes the *standard* gamma variate and callers apply the scale.
    """

    def __init__(self, seed: Optional[int] = None, rs: Optional[np.random.RandomState] = None):
        self.rs = rs if rs is not None else np.random.RandomState(seed)

    def reseed(self, seed: int) -> None:  # np.random.seed(seed) on the global stream (cluster.py:94-95)
        self.rs = np.random.RandomState(seed)

    def uniform_matrix(self, n: int, d: int) -> np.ndarray:  # np.random.rand(n, d)
        return self.rs.rand(n, d)

    def uniform_vector(self, n: int) -> np.ndarray:  # np.random.rand(n) / random_sample(n)
        return self.rs.random_sample(n)

    def uniform_scalar(self) -> float:  # np.random.random()
        return float(self.rs.random_sample())

    def normal_vector(self, d: int) -> np.ndarray:  # np.random.randn(d)
        return self.rs.randn(d)

    def standard_gamma(self, shape: float) -> float:
        return float(self.rs.standard_gamma(shape))

    def pick(self, pool: np.ndarray, size: int) -> np.ndarray:  # np.random.choice(pool, size)
        return self.rs.choice(pool, size=size, replace=True)


def legacy_choice_indices(p: np.ndarray, uniforms: np.ndarray) -> np.ndarray:
    """``np.random.choice(len(p), size, replace=True, p=p)`` given its uniforms.

    numpy's legacy RandomState.choice (third-party; verified equal on this image):
    ``cdf = p.cumsum(); cdf /= cdf[-1]; idx = cdf.searchsorted(U, side='right')``.
    Used by tempest/steps/resample.py:79-82 and tempest/modes.py:199-201,272-274.
    """
    cdf = np.cumsum(p)
    cdf /= cdf[-1]
    return cdf.searchsorted(uniforms, side="right")


# --------------------------------------------------------------------------------------
# Kernel (a): persistent-ensemble log-weights and evidence
# --------------------------------------------------------------------------------------
def log_weights_and_logz(
    logl_gens: Sequence[np.ndarray],
    beta_gens: Sequence[float],
    logz_gens: Sequence[float],
    beta_final: float = 1.0,
    normalize: bool = True,
    chunk: int = 1 << 18,
) -> Tuple[np.ndarray, float]:
    """Balance-heuristic MIS log-weights of every stored particle at ``beta_final``.

    Follows tempest/state_manager.py:418-480: for particle s and generations t,
    ``logw_s = beta_final*l_s - LSE_t(l_s*beta_t - logZ_t + log n_t - log N)`` with the
    LSE done by a *sequential* ``np.logaddexp.reduce`` in generation order (:466-473),
    ``logz = LSE_s(logw) - log N`` (:475) and optional normalisation (:477-478).
    Rows are processed in chunks only to bound memory; per-row arithmetic is unchanged.
    """
    beta = np.asarray(beta_gens, dtype=float)
    if beta.size == 0:
        return np.array([]), -np.inf  # :454-455
    logz_iter = np.asarray(logz_gens, dtype=float)
    logl_all = np.concatenate(logl_gens)
    n_per = np.array([len(g) for g in logl_gens])
    n_total = n_per.sum()
    log_mix = np.log(n_per) - np.log(n_total)
    big_b = np.empty_like(logl_all)
    for lo in range(0, logl_all.size, chunk):
        hi = min(lo + chunk, logl_all.size)
        b = logl_all[lo:hi, None] * beta[None, :] - logz_iter[None, :]
        big_b[lo:hi] = np.logaddexp.reduce(b + log_mix[None, :], axis=1)
    logw = logl_all * beta_final - big_b
    logz_new = np.logaddexp.reduce(logw) - np.log(logw.size)
    if normalize and logw.size:
        logw = logw - np.logaddexp.reduce(logw)
    return logw, float(logz_new)


def effective_sample_size(weights: np.ndarray) -> float:
    """``1 / sum((w/sum w)^2)`` -- tempest/tools.py:120-135."""
    w = weights / np.sum(weights)
    return float(1.0 / np.sum(w**2.0))


def volume_variation(x: np.ndarray, w: Optional[np.ndarray] = None) -> float:
    """Influence-function CV of sqrt(det Cov) -- tempest/tools.py:58-117."""
    x = np.asarray(x)
    n, d = x.shape
    if n < d + 1:
        return 1e10  # :87-88
    if w is None:
        w = np.ones(n)
    w = np.asarray(w)
    w = w / np.sum(w)
    mean = np.sum(x * w[:, None], axis=0)  # :96
    xc = x - mean
    cov = np.dot(xc.T, xc * w[:, None])  # :99
    if np.linalg.matrix_rank(cov) < d:  # :101-104
        cov = cov + np.eye(d) * (1e-6 * np.trace(cov))
    try:
        cov_inv = np.linalg.inv(cov)
    except np.linalg.LinAlgError:
        return 1e10
    d2 = np.sum(xc @ cov_inv * xc, axis=1)  # :111
    dev = np.clip(d2 - d, -1e6, 1e6)  # :114
    return float(0.5 * np.sqrt(np.sum(w**2 * dev**2)))  # :115


# --------------------------------------------------------------------------------------
# Kernel (b): next-beta search
# --------------------------------------------------------------------------------------
class BetaSearch:
    """ESS bracket + bisection of tempest/steps/reweight.py:88-297 over a probe callable.

    ``probe(beta) -> (weights_unnormalised, ess, metric)``; every probe is logged in
    ``self.log`` as ``(beta, ess)`` so the CUDA driver's probe sequence can be compared.
    """

    def __init__(self, probe: Callable[[float], Tuple[np.ndarray, float, float]], dynamic: bool):
        self._probe = probe
        self.dynamic = dynamic
        self.log: List[Tuple[float, float]] = []

    def probe(self, beta: float):
        out = self._probe(beta)
        self.log.append((float(beta), float(out[1])))
        return out

    def ess_bracket(self, beta_current: float, ess_target: float) -> Tuple[float, float]:
        """reweight.py:225-297."""
        lo, hi = beta_current, 1.0
        _, ess_cur, _ = self.probe(beta_current)
        if ess_cur <= ess_target:  # :264-266
            return beta_current, beta_current
        _, ess_one, _ = self.probe(1.0)
        if ess_one >= ess_target:  # :269-271
            return 1.0, 1.0
        while True:  # :277-295
            mid = (hi + lo) * 0.5
            scale = max(abs(lo), abs(hi), _TINY)
            if hi - lo <= max(BETA_RTOL * scale, BETA_TOLERANCE * scale):
                break
            _, ess_mid, _ = self.probe(mid)
            if ess_mid >= ess_target:
                lo = mid
            else:
                hi = mid
        return lo, hi

    def bisect(self, beta_min: float, beta_max: float, target: float, use_metric: bool):
        """reweight.py:123-223; returns (beta, weights, ess) of the LAST probe."""
        beta = beta_min
        w = ess = None
        for _ in range(MAX_BISECTION):
            beta = (beta_max + beta_min) * 0.5
            w, ess, metric = self.probe(beta)
            val = metric if use_metric else ess
            if not np.isfinite(val):
                val = 1e10  # :168-174
            atol = METRIC_ATOL_CV if self.dynamic else METRIC_ATOL
            metric_ok = abs(val - target) < max(ESS_TOLERANCE * abs(target), atol)
            scale = max(abs(beta_min), abs(beta_max), _TINY)
            beta_ok = (beta_max - beta_min) < max(BETA_RTOL * scale, BETA_TOLERANCE * scale)
            if metric_ok or beta_ok or beta == 1.0:  # :201
                return beta, w, ess
            if not use_metric:  # ESS falls as beta rises (:205-211)
                if val < target:
                    beta_max = beta
                else:
                    beta_min = beta
            else:  # CV rises with beta (:212-220)
                if val < target:
                    beta_min = beta
                else:
                    beta_max = beta
        return beta, w, ess


# --------------------------------------------------------------------------------------
# Trim / resampling (kernels c, e)
# --------------------------------------------------------------------------------------
def trim_weights(weights: np.ndarray, ess: float = TRIM_ESS, bins: int = TRIM_BINS):
    """tempest/tools.py:10-55.  NB normalises ``weights`` IN PLACE (:36); returns
    ``(kept_indices, trimmed_normalised_weights, bin_index_i)``."""
    weights /= np.sum(weights)
    ess_total = 1.0 / np.sum(weights**2.0)
    percentiles = np.linspace(0, 99, bins)
    i = bins - 1
    while True:
        thr = np.percentile(weights, percentiles[i])
        mask = weights >= thr
        wt = weights[mask]
        wt /= np.sum(wt)
        ess_trim = 1.0 / np.sum(wt**2.0)
        if ess_trim / ess_total >= ess:
            break
        i -= 1
    return np.nonzero(mask)[0], wt, i


def systematic_indices(size: int, weights: np.ndarray, u0: float) -> np.ndarray:
    """tempest/tools.py:178-228 given its single uniform ``u0`` (strict ``>`` walk of the
    sequential running sum, no final clamp)."""
    if abs(np.sum(weights) - 1.0) > SQRTEPS:
        weights = np.array(weights) / np.sum(weights)
    positions = (u0 + np.arange(size)) / size
    # sequential running sum == np.cumsum (strictly left-to-right)
    csum = np.cumsum(weights)
    # smallest j with NOT(pos > csum[j])  <=>  csum[j] >= pos  (searchsorted 'left')
    idx = np.searchsorted(csum, positions, side="left")
    if idx.size and idx.max() >= len(weights):
        raise IndexError("systematic walk ran past the last weight (reference would too)")
    return idx.astype(int)


def systematic_indices_loop(size: int, weights: np.ndarray, u0: float) -> np.ndarray:
    """Literal loop form of tools.py:217-226, for small cross-checks of the vector form."""
    if abs(np.sum(weights) - 1.0) > SQRTEPS:
        weights = np.array(weights) / np.sum(weights)
    positions = (u0 + np.arange(size)) / size
    j = 0
    run = weights[0]
    out = np.empty(size, dtype=int)
    for i in range(size):
        while positions[i] > run:
            j += 1
            run += weights[j]
        out[i] = j
    return out


# --------------------------------------------------------------------------------------
# Student-t fit and mode statistics (kernel e)
# --------------------------------------------------------------------------------------
def _func0_at_huge_nu(delta: np.ndarray, dim: int, n: int) -> float:
    """``func0(1e300)`` of tempest/student.py:41-54 (decides the nu=inf early exit)."""
    from scipy import special

    nu = 1e300
    w = (nu + dim) / (nu + delta)
    return float(
        -special.psi(nu / 2)
        + np.log(nu / 2)
        + np.sum(np.log(w)) / n
        - np.sum(w) / n
        + 1
        + special.psi((nu + dim) / 2)
        - np.log((nu + dim) / 2)
    )


def student_fit(data: np.ndarray, tolerance: float = 1e-6, max_iter: int = 100):
    """EM fit of a multivariate Student-t -- tempest/student.py:6-116.

    Returns ``(mu, Sigma, nu)``.  In practice the first E-step finds ``func0(1e300) >= 0``
    and returns ``(median, MLE cov + diag(var)/n, inf)``."""
    from scipy import optimize, special

    data = data.T
    dim, n = data.shape
    mu = np.array([np.median(data, 1)]).T  # :62
    sigma = np.cov(data) * (n - 1) / n + (1 / n) * np.diag(np.var(data, axis=1))  # :63
    sigma = np.atleast_2d(sigma)
    nu = 20
    last_nu = 0
    it = 0
    floor = 1e-6
    while abs(last_nu - nu) > tolerance and it < max_iter:
        it += 1
        try:
            np.linalg.cholesky(sigma)
        except np.linalg.LinAlgError:
            sigma = sigma + np.eye(dim) * max(floor, floor * abs(np.trace(sigma)))
        diffs = data - mu
        try:
            delta = np.sum(diffs * np.linalg.solve(sigma, diffs), 0)
        except np.linalg.LinAlgError:
            sigma = sigma + np.eye(dim) * max(1e-3, 1e-3 * abs(np.trace(sigma)))
            delta = np.sum(diffs * np.linalg.solve(sigma, diffs), 0)
        last_nu = nu
        if _func0_at_huge_nu(delta, dim, n) >= 0:  # :54-55
            return mu.T[0], sigma, np.inf  # :93-94

        def func0(v):
            w = (v + dim) / (v + delta)
            return (
                -special.psi(v / 2)
                + np.log(v / 2)
                + np.sum(np.log(w)) / n
                - np.sum(w) / n
                + 1
                + special.psi((v + dim) / 2)
                - np.log((v + dim) / 2)
            )

        nu = optimize.bisect(func0, 1e-300, 1e300)  # :57 (raises in practice)
        w = (nu + dim) / (nu + delta)
        sigma = np.dot(w * diffs, diffs.T) / n
        mu = np.array([np.sum(w * data, 1) / sum(w)]).T
    try:
        np.linalg.cholesky(sigma)
    except np.linalg.LinAlgError:
        sigma = sigma + np.eye(dim) * max(floor, floor * abs(np.trace(sigma)))
    return mu.T[0], sigma, nu


class ModeStats:
    """Per-mode mean/cov/dof with Cholesky factor and inverse -- tempest/modes.py:58-119."""

    def __init__(self, means, covs, dofs):
        self.means = np.atleast_2d(np.asarray(means, dtype=float))
        covs = np.asarray(covs, dtype=float)
        self.covs = covs.reshape(1, *covs.shape) if covs.ndim == 2 else covs.copy()
        self.dofs = np.atleast_1d(np.asarray(dofs, dtype=float))
        self.K, self.n_dim = self.means.shape
        self.inv = np.empty_like(self.covs)
        self.chol = np.empty_like(self.covs)
        for k in range(self.K):
            c = self.covs[k]
            try:
                self.chol[k] = np.linalg.cholesky(c)
                self.inv[k] = np.linalg.inv(c)
            except np.linalg.LinAlgError:  # :114-119
                c = c + np.eye(c.shape[0]) * max(1e-6, 1e-6 * abs(np.trace(c)))
                self.covs[k] = c
                self.chol[k] = np.linalg.cholesky(c)
                self.inv[k] = np.linalg.inv(c)


def mode_stats_global(u: np.ndarray, weights: np.ndarray, uniforms: np.ndarray,
                      factor: int = 4) -> Tuple[ModeStats, np.ndarray]:
    """Single global mode -- tempest/modes.py:221-288: 4n weighted draws, Student-t fit,
    non-finite dof -> DOF_FALLBACK.  ``uniforms`` are the 4n uniforms ``choice`` consumes."""
    w = weights / np.sum(weights)
    idx = legacy_choice_indices(w, uniforms)
    mean, cov, dof = student_fit(u[idx])
    if not np.isfinite(dof):
        dof = DOF_FALLBACK
    return ModeStats(mean.reshape(1, -1), cov.reshape(1, *cov.shape), np.array([dof])), idx


def mode_stats_particles(u: np.ndarray, weights: np.ndarray, labels: np.ndarray, stream,
                         factor: int = 4, record: Optional[dict] = None) -> ModeStats:
    """One Student-t mode per *distinct predicted label* -- tempest/modes.py:131-219.  Members keep
    their order (:191), weights are renormalised inside the cluster (:193-194), ``4 n_c`` rows are
    drawn with replacement (:197-201) and fitted (:204-209).  Mode k is the k-th distinct label
    (:188), so walkers carrying a label beyond K index out of range downstream."""
    w = weights / np.sum(weights)
    means, covs, dofs, unis, draws, members_all = [], [], [], [], [], []
    for label in np.unique(labels):
        members = np.where(labels == label)[0]
        wc = w[members]
        wc = wc / np.sum(wc)
        uni = stream.uniform_vector(factor * len(members))
        idx = legacy_choice_indices(wc, uni)
        mean, cov, dof = student_fit(u[members][idx])
        if not np.isfinite(dof):
            dof = DOF_FALLBACK
        means.append(mean)
        covs.append(cov)
        dofs.append(dof)
        unis.append(uni)
        draws.append(idx)
        members_all.append(members)
    if record is not None:
        record.update(train_u=np.concatenate(unis), draw_idx=draws, members=members_all)
    return ModeStats(np.array(means), np.array(covs), np.array(dofs))


# --------------------------------------------------------------------------------------
# Boundaries (mcmc.py:326-411)
# --------------------------------------------------------------------------------------
def boundary_map(u: np.ndarray, periodic=None, reflective=None) -> np.ndarray:
    """Periodic wrap (``% 1``) and reflective fold (floor-parity flip) -- mcmc.py:326-366."""
    out = np.array(u, dtype=float, copy=True)
    if periodic is not None:
        for j in periodic:
            out[..., j] = out[..., j] % 1.0
    if reflective is not None:
        for j in reflective:
            v = out[..., j]
            k = np.floor(v).astype(int)
            r = v - k
            out[..., j] = np.where(k % 2 == 0, r, 1.0 - r)
    return out


def inside_unit_cube(u: np.ndarray, periodic=None, reflective=None):
    """True when every non-periodic, non-reflective coordinate is in [0,1] -- mcmc.py:369-411."""
    d = u.shape[-1]
    special = set()
    if periodic is not None:
        special.update(int(j) for j in periodic)
    if reflective is not None:
        special.update(int(j) for j in reflective)
    strict = [j for j in range(d) if j not in special]
    if not strict:
        return True if u.ndim == 1 else np.ones(u.shape[0], dtype=bool)
    s = u[..., strict]
    if u.ndim == 1:
        return bool(np.all(s >= 0) and np.all(s <= 1))
    return np.all(s >= 0, axis=-1) & np.all(s <= 1, axis=-1)


# --------------------------------------------------------------------------------------
# Kernel (d): MCMC mutation
# --------------------------------------------------------------------------------------
def adaptive_steps(n_steps, n_max, n_dim, sigma_0, sigmas, assignments, n_clusters, acc):
    """mcmc.py:104-135 (note ``sigmas[:n_nonempty]`` indexing)."""
    sizes = np.array([np.sum(assignments == c) for c in range(n_clusters)])
    sizes = sizes[sizes > 0]
    wsig = np.average(sigmas[: len(sizes)], weights=sizes)
    n_min = n_steps * n_dim
    n_adapt = n_steps * n_dim * (0.234 / max(0.01, acc)) * (sigma_0 / max(1e-6, wsig)) ** 2
    return int(min(max(n_min, n_adapt), n_max * n_dim))


def mcmc_mutate(
    u, x, logl, assignments, beta, stats: ModeStats, log_likelihood, prior_transform,
    stream, n_steps, n_max, sample="tpcn", periodic=None, reflective=None, record=None,
):
    """t-preconditioned Crank-Nicolson / random-walk Metropolis -- tempest/mcmc.py:142-323.

    Walker-by-walker proposals (gamma then randn(D), redrawn until inside the cube, :225-249 /
    :301-312), one batched likelihood call (:157-160), ``alpha = nan_to_num(min(1, exp(beta*
    (l'-l) + factor)))`` (:163-166), one uniform per walker (:169-170), per-cluster Robbins-Monro
    sigma update (:180-186, 281-288 / 320-323) and the adaptive stop (:192-194).
    ``record`` (dict) receives the variates and the per-step diagnostics."""
    u = u.copy()
    x = x.copy()
    logl = logl.copy()
    n, d = x.shape
    K = stats.K
    tpcn = sample != "rwm"
    sigma_0 = 2.38 / np.sqrt(d)
    sigmas = np.ones(K) * (np.minimum(sigma_0, 0.99) if tpcn else sigma_0)
    it = 0
    calls = 0
    if record is not None:
        record.update(gamma=[], z=[], acc_u=[], sigma=[], alpha=[], accept=[], u_prop=[])
    while True:
        it += 1
        u_prop = np.empty_like(u)
        g_row = np.zeros(n)
        z_row = []
        for k in range(n):
            c = assignments[k]
            sig = sigmas[c]
            chol = stats.chol[c]
            if tpcn:
                mu = stats.means[c]
                diff = u[k] - mu
                q = diff @ stats.inv[c] @ diff
                g = stream.standard_gamma((d + stats.dofs[c]) / 2)
                g_row[k] = g
                s = 1.0 / ((2.0 / (stats.dofs[c] + q)) * g)  # 1/np.random.gamma(shape, scale)
            zs = []
            while True:
                z = stream.normal_vector(d)
                zs.append(z)
                if tpcn:
                    prop = mu + np.sqrt(1.0 - sig**2.0) * diff + sig * np.sqrt(s) * chol @ z
                else:
                    prop = u[k] + sig * chol @ z
                prop = boundary_map(prop, periodic, reflective)
                if inside_unit_cube(prop, periodic, reflective):
                    break
            u_prop[k] = prop
            z_row.append(np.array(zs))
        x_prop = np.array([prior_transform(r) for r in u_prop])
        logl_prop = np.asarray(log_likelihood(x_prop))
        calls += n
        if tpcn:  # mcmc.py:251-279
            mus = stats.means[assignments]
            dof = stats.dofs[assignments]
            inv = stats.inv[assignments]
            dcur = u - mus
            qc = np.einsum("ij,ijk,ik->i", dcur, inv, dcur)
            B = -0.5 * (d + dof) * np.log(1 + qc / dof)
            dnew = u_prop - mus
            qn = np.einsum("ij,ijk,ik->i", dnew, inv, dnew)
            A = -0.5 * (d + dof) * np.log(1 + qn / dof)
            factor = -A + B
        else:
            factor = np.zeros(n)
        with np.errstate(over="ignore", invalid="ignore"):
            alpha = np.exp(beta * (logl_prop - logl) + factor)
        alpha = np.minimum(1.0, alpha)
        alpha = np.nan_to_num(alpha, nan=0.0)
        u_rand = stream.uniform_vector(n)
        accept = u_rand < alpha
        if record is not None:
            record["gamma"].append(g_row)
            record["z"].append(z_row)
            record["acc_u"].append(u_rand)
            record["sigma"].append(sigmas.copy())
            record["alpha"].append(alpha.copy())
            record["accept"].append(accept.copy())
            record["u_prop"].append(u_prop.copy())
        u[accept] = u_prop[accept]
        x[accept] = x_prop[accept]
        logl[accept] = logl_prop[accept]
        for c in range(K):
            m = assignments == c
            if not np.any(m):
                continue
            mean_alpha = alpha[m].mean()
            rate = 1.0 / (it + 1)
            if tpcn:
                sigmas[c] = np.clip(sigmas[c] + rate * (mean_alpha - 0.234), 0, min(sigma_0, 0.99))
            else:
                sigmas[c] = sigmas[c] + rate * (mean_alpha - 0.234)
        acc_now = accept.mean()
        if it >= adaptive_steps(n_steps, n_max, d, sigma_0, sigmas, assignments, K, acc_now):
            break
    return u, x, logl, float(sigmas.mean() / sigma_0), float(alpha.mean()), it, calls


# --------------------------------------------------------------------------------------
# Full PS loop (core.py:110-185, 360-374; steps/*.py)
# --------------------------------------------------------------------------------------
class OraclePS:
    """Persistent Sampling restated end to end (``clustering`` on or off).

    ``iterate()`` is ``SamplerCore.execute_iteration`` (core.py:162-185): reweight
    (steps/reweight.py:341-495) -> train (steps/train.py:65-127, global branch) -> resample
    (steps/resample.py:52-99) -> mutate (steps/mutate.py:76-200) -> commit
    (state_manager.py:356-416).  ``run()`` is ``run_sampling`` (core.py:110-160)."""

    def __init__(
        self, prior_transform, log_likelihood, n_dim, n_particles=None, ess_ratio=2.0,
        volume_variation=None, periodic=None, reflective=None, sample="tpcn", n_steps=None,
        n_max_steps=None, resample="mult", stream=None, record=False, clustering=False,
        normalize=True, cluster_every=1, split_threshold=1.0, n_max_clusters=None,
    ):
        self.prior_transform = prior_transform
        self.log_likelihood = log_likelihood
        self.n_dim = int(n_dim)
        self.n_particles = int(n_particles) if n_particles is not None else 2 * self.n_dim
        self.ess_ratio = ess_ratio
        self.volume_variation = volume_variation
        self.periodic = periodic
        self.reflective = reflective
        self.sample = sample
        self.n_steps = 1 if (n_steps is None or n_steps <= 0) else int(n_steps)  # config.py:80-81
        self.n_max_steps = (20 * self.n_steps if (n_max_steps is None or n_max_steps <= 0)
                            else int(n_max_steps))  # config.py:83-84
        self.resample = resample
        self.stream = stream if stream is not None else LegacyStream(0)
        self.record = record
        self.clustering = bool(clustering)
        self.normalize = bool(normalize)
        self.cluster_every = int(cluster_every)
        self.split_threshold = float(split_threshold)
        self.n_max_clusters = n_max_clusters
        self.clusterer = None
        self.hist: Dict[str, list] = {k: [] for k in (
            "u", "x", "logl", "iter", "logz", "calls", "steps", "efficiency", "ess", "cv",
            "acceptance", "beta")}
        self.cur: Dict[str, object] = dict(iter=0, calls=0, beta=0.0, logz=0.0)
        self.tapes: List[dict] = []
        self.traces: List[dict] = []
        self.n_total = 0

    # -- state helpers -----------------------------------------------------------------
    def logw_logz(self, beta_final=1.0):
        return log_weights_and_logz(self.hist["logl"], self.hist["beta"], self.hist["logz"], beta_final)

    def _probe(self, beta):
        logw, _ = self.logw_logz(beta)
        w = np.exp(logw - np.max(logw))
        ess = effective_sample_size(w)
        metric = ess
        if self.volume_variation is not None:
            metric = volume_variation(np.concatenate(self.hist["u"]), w / np.sum(w))
        return w, ess, metric

    # -- steps -------------------------------------------------------------------------
    def _reweight(self, trace):
        self.cur["iter"] = self.cur["iter"] + 1
        n = self.n_particles
        if len(self.hist["beta"]) == 0:  # reweight.py:365-383
            self.cur.update(beta=0.0, logz=0.0, ess=self.ess_ratio * n, cv=0.0)
            trace["probes"] = []
            return np.ones(n) / n
        dynamic = self.volume_variation is not None
        search = BetaSearch(self._probe, dynamic)
        beta_prev = self.cur["beta"]
        target = self.ess_ratio * n
        lo, hi = search.ess_bracket(beta_prev, target)
        if lo == hi:
            beta = lo
            w, ess, _ = search.probe(beta)
        elif not dynamic:
            beta, w, ess = search.bisect(beta_prev, hi, target, use_metric=False)
        else:  # reweight.py:427-482
            _, ess_prev, cv_prev = search.probe(beta_prev)
            _, ess_high, cv_high = search.probe(hi)
            if self.volume_variation >= cv_high:
                beta, w, ess = hi, None, ess_high
            elif self.volume_variation <= cv_prev:
                beta, w, ess = beta_prev, None, ess_prev
            else:
                beta, w, ess = search.bisect(beta_prev, hi, self.volume_variation, use_metric=True)
            if w is None:
                w, ess, _ = search.probe(beta)
        cv = volume_variation(np.concatenate(self.hist["u"]), w / np.sum(w))
        _, logz = self.logw_logz(beta)
        self.cur.update(logz=logz, beta=beta, ess=ess, cv=cv)
        trace["probes"] = list(search.log)
        return w / np.sum(w)

    def _train(self, weights, tape, trace):
        d = self.n_dim
        if self.cur["beta"] == 0.0:  # train.py:79-88
            return ModeStats(np.zeros((1, d)), np.eye(d).reshape(1, d, d), np.array([DOF_FALLBACK]))
        idx, w_trim, i_bin = trim_weights(weights)
        u = np.concatenate(self.hist["u"])[idx]
        if self.clustering:  # train.py:97-115
            from .cluster_oracle import fit_hierarchy

            it = self.cur["iter"]
            if it % self.cluster_every == 0 or it == 0:
                cap = self.n_max_clusters  # core.py:59-69
                self.clusterer = fit_hierarchy(
                    u, w_trim, self.stream, normalize=self.normalize,
                    max_iterations=1000 if cap is None else cap - 1,
                    min_points=None if cap is None else 4 * d,
                    threshold_modifier=self.split_threshold)
            labels = self.clusterer.predict(u)
            rec: dict = {}
            stats = mode_stats_particles(u, w_trim, labels, self.stream, record=rec)
            tape["train_u"] = rec["train_u"]
            trace.update(trim_idx=idx, trim_w=w_trim, trim_bin=i_bin, train_labels=labels,
                         train_draw_idx=rec["draw_idx"], n_clusters=self.clusterer.n_clusters,
                         cluster_centres=np.array(self.clusterer.centres),
                         cluster_covs=np.array(self.clusterer.covs),
                         cluster_weights=self.clusterer.weights.copy(),
                         mode_mean=stats.means.copy(), mode_cov=stats.covs.copy(),
                         mode_chol=stats.chol.copy(), mode_inv=stats.inv.copy(), mode_dof=stats.dofs.copy())
            return stats
        uni = self.stream.uniform_vector(4 * len(idx))
        stats, draw_idx = mode_stats_global(u, w_trim, uni)
        tape["train_u"] = uni
        trace.update(trim_idx=idx, trim_w=w_trim, trim_bin=i_bin, train_draw_idx=draw_idx,
                     mode_mean=stats.means.copy(), mode_cov=stats.covs.copy(),
                     mode_chol=stats.chol.copy(), mode_inv=stats.inv.copy(), mode_dof=stats.dofs.copy())
        return stats

    def _resample(self, weights, tape, trace):
        n = self.n_particles
        if self.cur["beta"] == 0.0:  # resample.py:69-72
            self.cur["assignments"] = np.zeros(n, dtype=int)
            return
        u = np.concatenate(self.hist["u"])
        x = np.concatenate(self.hist["x"])
        logl = np.concatenate(self.hist["logl"])
        if self.resample == "mult":
            uni = self.stream.uniform_vector(n)
            idx = legacy_choice_indices(weights, uni)
            tape["resample_u"] = uni
        else:
            u0 = self.stream.uniform_scalar()
            idx = systematic_indices(n, weights, u0)
            tape["resample_u"] = np.array([u0])
        trace["resample_idx"] = idx
        trace["resample_p"] = weights.copy()
        assign = self.clusterer.predict(u[idx]) if self.clustering else np.zeros(n, dtype=int)  # :92-94
        trace["assignments"] = assign
        self.cur.update(u=u[idx], x=x[idx], logl=logl[idx], assignments=assign)

    def _mutate(self, stats, tape, trace):
        n, d = self.n_particles, self.n_dim
        if self.cur["beta"] == 0.0:  # mutate.py:100-149
            u = self.stream.uniform_matrix(n, d)
            tape["prior_u"] = u.copy()
            x = np.array([self.prior_transform(u[i]) for i in range(n)])
            logl = np.asarray(self.log_likelihood(x), dtype=float)
            self.cur.update(u=u, x=x, logl=logl, assignments=np.zeros(n, dtype=int),
                            calls=self.cur["calls"] + n, steps=1, acceptance=1.0, efficiency=1.0)
            bad = np.isinf(logl)
            if np.any(bad):
                every = np.arange(n)
                inf_idx, fin_idx = every[bad], every[~bad]
                if len(fin_idx) > 0:
                    pick = self.stream.pick(fin_idx, len(inf_idx))
                    tape["inf_pick"] = pick
                    x[inf_idx] = x[pick]
                    u[inf_idx] = u[pick]
                    logl[inf_idx] = logl[pick]
                self.cur["logz"] = self.cur["logz"] + np.log(len(fin_idx) / n)
            return
        rec = {} if self.record else None
        u, x, logl, eff, acc, steps, calls = mcmc_mutate(
            self.cur["u"], self.cur["x"], self.cur["logl"], self.cur["assignments"],
            self.cur["beta"], stats, self.log_likelihood, self.prior_transform, self.stream,
            self.n_steps, self.n_max_steps, self.sample, self.periodic, self.reflective, rec)
        if rec is not None:
            tape.update(gamma=rec["gamma"], z=rec["z"], acc_u=rec["acc_u"])
            trace.update(mcmc_sigma=rec["sigma"], mcmc_alpha=rec["alpha"], mcmc_accept=rec["accept"],
                         mcmc_u_prop=rec["u_prop"])
        self.cur.update(u=u, x=x, logl=logl, efficiency=eff, acceptance=acc, steps=steps,
                        calls=self.cur["calls"] + calls)

    def iterate(self) -> dict:
        tape: dict = {}
        trace: dict = {}
        weights = self._reweight(trace)
        trace["weights"] = weights.copy()
        stats = self._train(weights, tape, trace)
        self._resample(weights, tape, trace)
        self._mutate(stats, tape, trace)
        for k in self.hist:  # commit (state_manager.py:410-415)
            v = self.cur.get(k)
            if v is not None:
                self.hist[k].append(np.copy(v) if isinstance(v, np.ndarray) else v)
        trace.update({k: self.cur.get(k) for k in ("iter", "beta", "logz", "ess", "cv", "steps",
                                                     "acceptance", "efficiency", "calls")})
        trace.update(u=self.cur["u"].copy(), x=self.cur["x"].copy(), logl=self.cur["logl"].copy())
        self.tapes.append(tape)
        self.traces.append(trace)
        return dict(self.cur)

    def not_terminated(self) -> bool:
        """core.py:360-374."""
        logw, _ = self.logw_logz(1.0)
        if len(logw) == 0:
            return True
        ess = effective_sample_size(np.exp(logw - np.max(logw)))
        return 1.0 - self.cur["beta"] >= 1e-4 or ess < self.n_total

    def run(self, n_total: int = 4096, max_iterations: Optional[int] = None):
        """core.py:110-160 (fresh run; history is not cleared, :376-381)."""
        self.cur.update(iter=0, calls=0, beta=0.0, logz=0.0)
        self.n_total = int(n_total)
        k = 0
        while self.not_terminated():
            self.iterate()
            k += 1
            if max_iterations is not None and k >= max_iterations:
                break
        _, logz = self.logw_logz(1.0)
        self.cur["logz"] = logz

    def posterior(self, resample=False, trim_importance_weights=True, return_logw=False,
                  ess_trim=0.99, bins_trim=1000):
        """core.py:187-242."""
        logw, _ = self.logw_logz(1.0)
        w = np.exp(logw - np.max(logw))
        w /= np.sum(w)
        x = np.concatenate(self.hist["x"])
        logl = np.concatenate(self.hist["logl"])
        if trim_importance_weights:
            idx, w, _ = trim_weights(w, ess_trim, bins_trim)
            x, logl = x[idx], logl[idx]
        if resample:
            idx = systematic_indices(len(w), w, self.stream.uniform_scalar())
            x, logl = x[idx], logl[idx]
            w = np.ones(len(idx)) / len(idx)
        return (x, w, logl, logw) if return_logw else (x, w, logl)

    def evidence(self):
        return self.cur["logz"], None


# --------------------------------------------------------------------------------------
# numpy pairwise summation restated (used to pin the all-equal-weights warm-up ESS)
# --------------------------------------------------------------------------------------
def numpy_pairwise_sum(a: Sequence[float]) -> float:
    """numpy's ``DOUBLE_pairwise_sum`` (third-party, numpy/_core/src/umath/loops_utils.h):
    n<8 sequential from 0; n<=128 eight running lanes combined as ((0+1)+(2+3))+((4+5)+(6+7))
    then the tail sequentially; otherwise split at ``n/2 - (n/2)%8``.  Verified == np.sum."""
    n = len(a)
    if n < 8:
        r = 0.0
        for v in a:
            r += v
        return r
    if n <= 128:
        lanes = [a[j] for j in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                lanes[j] += a[i + j]
            i += 8
        res = ((lanes[0] + lanes[1]) + (lanes[2] + lanes[3])) + ((lanes[4] + lanes[5]) + (lanes[6] + lanes[7]))
        while i < n:
            res += a[i]
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return numpy_pairwise_sum(a[:n2]) + numpy_pairwise_sum(a[n2:])
