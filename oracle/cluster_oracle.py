"""CPU oracle for the clustering row of the hot path: the weighted
Gaussian mixture and the divisive BIC-gated hierarchy of tempest/cluster.py, restated with numpy.

TEST INFRASTRUCTURE ONLY.  Nothing under ``tempest_b200/`` imports this module; it is the checker
for ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs.  Every function cites the
reference lines it follows.  Pinned against the reference itself by ``oracle/gen_golden.py``
(fixtures ``tests/golden/cluster_*.npz``, checked in ``tests/test_oracle_golden.py``).

Third-party arithmetic restated here: ``scipy.stats.multivariate_normal.logpdf`` (scipy 1.18.1 in
this image; ``_multivariate.py`` ``_PSD`` / ``_logpdf``): eigendecomposition of the covariance,
``eps = 1e6 * DBL_EPSILON * max|lambda|``; a negative eigenvalue below ``-eps`` raises ValueError,
any eigenvalue ``<= eps`` raises LinAlgError (``allow_singular=False``);
``logpdf = -0.5 * (D log 2pi + sum log lambda + |((x - mu) V) / sqrt(lambda)|^2)``.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

LOG_2PI = float(np.log(2.0 * np.pi))
REG_COVAR = 1e-6          # cluster.py:37
EM_TOL = 1e-3             # cluster.py:36
EM_MAX_ITER = 1000        # cluster.py:34
GMM_SEED = 42             # cluster.py:466,475,536
RESP_EPS = 1e-10          # cluster.py:191,208,225,282


class SingularCovariance(np.linalg.LinAlgError):
    pass


def mvn_whitener(cov: np.ndarray) -> Tuple[np.ndarray, float]:
    """scipy ``_PSD(cov, allow_singular=False)``: returns ``(U, log_pdet)`` with
    ``U = V / sqrt(lambda)`` so that the Mahalanobis distance is ``|dev @ U|^2``."""
    from scipy import linalg

    lam, vec = linalg.eigh(cov, lower=True, check_finite=True)
    eps = 1e6 * np.finfo(lam.dtype).eps * np.max(np.abs(lam))
    if np.min(lam) < -eps:
        raise ValueError("the input matrix must be symmetric positive semidefinite")
    if np.any(lam <= eps):
        raise SingularCovariance("singular matrix")
    return vec * np.sqrt(1.0 / lam), float(np.sum(np.log(lam)))


def mvn_logpdf(x: np.ndarray, mean: np.ndarray, cov: np.ndarray) -> np.ndarray:
    """``multivariate_normal.logpdf(x, mean, cov)`` for a full-rank covariance."""
    whit, log_pdet = mvn_whitener(np.asarray(cov, dtype=float))
    dev = np.asarray(x, dtype=float) - mean
    maha = np.sum(np.square(dev @ whit), axis=-1)
    return -0.5 * (cov.shape[0] * LOG_2PI + log_pdet + maha)


class MixtureFit:
    """Result of one weighted EM fit (cluster.py:56-133): ``weights[K]``, ``means[K,D]``,
    ``covs[K,D,D]`` (covariance_type='full'), ``n_iter`` and the lower bound of the step before
    the last (what the reference stores in ``lower_bound_``)."""

    def __init__(self, weights, means, covs, n_iter, lower_bound):
        self.weights, self.means, self.covs = weights, means, covs
        self.n_iter, self.lower_bound = n_iter, lower_bound
        self.K = len(weights)

    # cluster.py:264-283
    def bound(self, x: np.ndarray, sample_weight: np.ndarray) -> float:
        return mixture_lower_bound(x, self.weights, self.means, self.covs, sample_weight)

    # cluster.py:310-340 (full covariances)
    def bic(self, x: np.ndarray) -> float:
        n, d = x.shape
        n_par = (self.K - 1) + self.K * d + self.K * d * (d + 1) / 2
        ll = self.bound(x, np.ones(n) / n) * n
        return -2 * ll + n_par * np.log(n)

    # cluster.py:285-308
    def predict(self, x: np.ndarray) -> np.ndarray:
        n, d = x.shape
        lp = np.zeros((n, self.K))
        for k in range(self.K):
            try:
                lp[:, k] = np.log(self.weights[k] + RESP_EPS) + mvn_logpdf(
                    x, self.means[k], self.covs[k] + np.eye(d) * REG_COVAR)
            except (np.linalg.LinAlgError, ValueError):
                lp[:, k] = -np.inf
        return np.argmax(lp, axis=1)


def mixture_densities(x, weights, means, covs) -> np.ndarray:
    """``weights[k] * N(x; mu_k, Sigma_k + reg I)`` per component -- cluster.py:172-188."""
    n, d = x.shape
    dens = np.zeros((n, len(weights)))
    for k in range(len(weights)):
        try:
            dens[:, k] = weights[k] * np.exp(mvn_logpdf(x, means[k], covs[k] + np.eye(d) * REG_COVAR))
        except (np.linalg.LinAlgError, ValueError):     # :185-188 falls back to reg*I
            dens[:, k] = weights[k] * np.exp(mvn_logpdf(x, means[k], np.eye(d) * REG_COVAR))
    return dens


def mixture_lower_bound(x, weights, means, covs, sample_weight) -> float:
    """cluster.py:264-283: ``sum_i sw_i log(sum_k w_k N_k(x_i) + 1e-10)``; a component whose
    covariance scipy rejects is skipped (:278-279)."""
    n, d = x.shape
    tot = np.zeros(n)
    for k in range(len(weights)):
        try:
            tot += weights[k] * np.exp(mvn_logpdf(x, means[k], covs[k] + np.eye(d) * REG_COVAR))
        except (np.linalg.LinAlgError, ValueError):
            pass
    return float(np.sum(sample_weight * np.log(tot + RESP_EPS)))


def mixture_m_step(x, resp, sample_weight):
    """cluster.py:195-250 (full covariances)."""
    wr = resp * sample_weight[:, None]
    mass = np.sum(wr, axis=0)
    weights = mass / np.sum(mass)
    means = (wr.T @ x) / (mass[:, None] + RESP_EPS)
    d = x.shape[1]
    covs = np.zeros((resp.shape[1], d, d))
    for k in range(resp.shape[1]):
        diff = x - means[k]
        covs[k] = (wr[:, k] * diff.T) @ diff
        covs[k] /= np.sum(wr[:, k]) + RESP_EPS
    return weights, means, covs


def kmeanspp_centres(x, sample_weight, n_components, stream, picks: Optional[list] = None):
    """Weighted k-means++ -- cluster.py:139-158.  One ``np.random.rand()`` per centre; note the
    *left* ``searchsorted`` (numpy default) on the running sums."""
    centres = np.zeros((n_components, x.shape[1]))
    run = np.cumsum(sample_weight)
    j = int(np.searchsorted(run, stream.uniform_scalar() * run[-1]))
    centres[0] = x[j]
    chosen = [j]
    for k in range(1, n_components):
        dist = np.min([np.sum((x - centres[c]) ** 2, axis=1) for c in range(k)], axis=0)
        p = dist * sample_weight
        p /= np.sum(p)
        run = np.cumsum(p)
        j = int(np.searchsorted(run, stream.uniform_scalar() * run[-1]))
        centres[k] = x[j]
        chosen.append(j)
    if picks is not None:
        picks.extend(chosen)
    return centres


def fit_mixture(x: np.ndarray, sample_weight: np.ndarray, n_components: int, stream,
                max_iter: int = EM_MAX_ITER, tol: float = EM_TOL, log: Optional[dict] = None) -> MixtureFit:
    """``GaussianMixture(n_components, 'full', n_init=1, random_state=42).fit`` -- cluster.py:56-133.

    ``stream.reseed(42)`` is the reference's ``np.random.seed(self.random_state)`` (:94-95): it
    resets the *global* legacy stream, so everything drawn after a fit continues from there."""
    x = np.asarray(x, dtype=float)
    sw = np.asarray(sample_weight, dtype=float)
    sw = sw / np.sum(sw)                                   # :88
    stream.reseed(GMM_SEED)                                # :94-95
    picks: list = []
    centres = kmeanspp_centres(x, sw, n_components, stream, picks)
    resp = np.zeros((x.shape[0], n_components))            # :161-165
    for k in range(n_components):
        resp[:, k] = np.exp(-0.5 * np.sum((x - centres[k]) ** 2, axis=1))
    resp /= np.sum(resp, axis=1, keepdims=True)
    weights, means, covs = mixture_m_step(x, resp, sw)     # :168
    lower = -np.inf
    bounds = []
    it = -1
    for it in range(max_iter):                             # :103-121
        dens = mixture_densities(x, weights, means, covs)
        resp = dens / (np.sum(dens, axis=1, keepdims=True) + RESP_EPS)   # :191
        weights, means, covs = mixture_m_step(x, resp, sw)
        new_lower = mixture_lower_bound(x, weights, means, covs, sw)
        bounds.append(new_lower)
        if new_lower - lower < tol:
            break
        lower = new_lower
    if log is not None:
        log.update(picks=picks, bounds=bounds)
    return MixtureFit(weights, means, covs, it + 1, lower)


class HierarchyFit:
    """Fitted ``HierarchicalGaussianMixture`` (cluster.py:343-572), covariance_type='full'."""

    def __init__(self):
        self.labels = None
        self.centres: List[np.ndarray] = []
        self.covs: List[np.ndarray] = []
        self.weights = None
        self.n_clusters = 0
        self.data_min = None
        self.data_max = None
        self.normalize = False
        self.rounds: List[dict] = []          # diagnostics: per split round, per cluster

    def _norm(self, x):                                    # :377-383
        return (x - self.data_min) / (self.data_max - self.data_min + 1e-10)

    def predict(self, x: np.ndarray) -> np.ndarray:
        """cluster.py:574-600 + 633-696 (the argmax of the normalised probabilities equals the
        argmax of ``log N_k + log(w_k + 1e-10)``)."""
        return np.argmax(self.log_scores(x), axis=1)

    def log_scores(self, x: np.ndarray) -> np.ndarray:
        x = np.asarray(x, dtype=float)
        if self.normalize:
            x = self._norm(x)
        d = x.shape[1]
        out = np.zeros((x.shape[0], self.n_clusters))
        for k in range(self.n_clusters):
            if self.normalize:                             # :645-651
                mean = self._norm(self.centres[k])
                scale = self.data_max - self.data_min
                cov = self.covs[k] / np.outer(scale, scale)
            else:
                mean, cov = self.centres[k], self.covs[k]
            try:
                lp = mvn_logpdf(x, mean, cov + np.eye(d) * 1e-6)          # :663-666
            except Exception:
                lp = mvn_logpdf(x, mean, np.eye(d))                      # :667-670
            out[:, k] = lp + np.log(self.weights[k] + 1e-10)             # :691
        return out


def fit_hierarchy(x: np.ndarray, sample_weight: np.ndarray, stream, normalize: bool = True,
                  max_iterations: int = 1000, min_points: Optional[int] = None,
                  threshold_modifier: float = 1.0) -> HierarchyFit:
    """``HierarchicalGaussianMixture.fit`` -- cluster.py:420-572.

    Each round refits a 1- and a 2-component mixture to every cluster with at least
    ``min_points`` members (:452-478), keeps the split with the largest BIC improvement above
    ``threshold_modifier * (D + D(D+1)/2 + 1) * log ESS(weights)`` (:413-418, 493) whose children
    both keep ``min_points`` members (:494-501), and stops when no split qualifies (:503-506)."""
    fit = HierarchyFit()
    fit.normalize = bool(normalize)
    x = np.asarray(x, dtype=float)
    sw = np.asarray(sample_weight, dtype=float)
    n, d = x.shape
    if normalize:                                          # :436-439
        fit.data_min = np.min(x, axis=0)
        fit.data_max = np.max(x, axis=0)
        x = fit._norm(x)
    need = min_points if min_points is not None else 2 * d  # :441
    clusters: List[np.ndarray] = [np.arange(n)]
    rounds = 0
    while rounds < max_iterations:                         # :445
        rounds += 1
        best_gain, best_split, best_parent = -np.inf, None, None
        diag = []
        for c, members in enumerate(clusters):
            if len(members) < need:
                diag.append(None)
                continue
            data, w = x[members], sw[members]
            wn = w / np.sum(w)
            ess = 1.0 / np.sum(wn ** 2)                    # :407-411
            threshold = threshold_modifier * ((d + d * (d + 1) / 2 + 1) * np.log(ess))   # :413-418,460
            parent = fit_mixture(data, w, 1, stream)
            parent_bic = parent.bic(data)
            child = fit_mixture(data, w, 2, stream)
            child_bic = child.bic(data)
            gain = parent_bic - child_bic
            diag.append(dict(size=len(members), parent_bic=parent_bic, child_bic=child_bic, gain=gain,
                             threshold=threshold, parent_iter=parent.n_iter, child_iter=child.n_iter))
            if gain > threshold and gain > best_gain:      # :493
                side = child.predict(data)
                a, b = members[side == 0], members[side == 1]
                if len(a) >= need and len(b) >= need:      # :497
                    best_gain, best_split, best_parent = gain, (a, b), c
        fit.rounds.append(dict(clusters=diag, split=best_parent))
        if best_split is None:
            break
        clusters.pop(best_parent)                          # :508-509
        clusters.extend(best_split)
    labels = np.full(n, -1, dtype=int)
    for c, members in enumerate(clusters):                 # :527-556
        data, w = x[members], sw[members]
        if len(data) >= d:
            one = fit_mixture(data, w, 1, stream)
            centre, cov = one.means[0], one.covs[0]
        else:
            centre, cov = np.mean(data, axis=0), np.eye(d)
        if normalize:                                      # :385-405
            scale = fit.data_max - fit.data_min
            centre = centre * scale + fit.data_min
            cov = cov * np.outer(scale, scale)
        fit.centres.append(centre)
        fit.covs.append(cov)
        labels[members] = c
    fit.labels = labels
    fit.n_clusters = len(clusters)
    total = np.sum(sw)                                     # :563-569
    fit.weights = np.array([np.sum(sw[labels == c]) / total for c in range(fit.n_clusters)])
    return fit
