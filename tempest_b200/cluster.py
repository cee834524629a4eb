"""Device mirror of ``tempest/cluster.py``: weighted Gaussian mixtures fitted by EM and the divisive,
BIC-gated ``HierarchicalGaussianMixture`` the reference uses to find modes (cluster.py:343-696).

All passes over the data (k-means++ probabilities and picks, E-steps, M-step moments, bounds, label
prediction, min-max normalisation, member-list splits) are CUDA kernels behind the C ABI
(csrc/tb_cluster.cu); the host keeps only what the reference keeps in Python control flow: which
cluster to split next (cluster.py:445-521) and the final denormalisation of K small matrices.
Every ``GaussianMixture.fit`` of the reference reseeds numpy with 42 (cluster.py:94-95,466,475,536),
so its k-means++ uniforms are the first values of MT19937(42): they are constants here.
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np
import torch

from . import _lib
from .ensemble import ptr, stream_ptr

F64 = torch.float64
REG_COVAR = 1e-6          # cluster.py:37
EM_TOL = 1e-3             # cluster.py:36
EM_MAX_ITER = 1000        # cluster.py:34
GMM_SEED = 42             # cluster.py:466,475,536
FIRST_CHUNK = 3           # EM passes enqueued before the first look at the device-side stop flag
NEXT_CHUNK = 8


def _seed_uniforms(n: int) -> np.ndarray:
    """The ``np.random.rand()`` values every reference fit sees after ``np.random.seed(42)``."""
    return np.random.RandomState(GMM_SEED).random_sample(n)


class MixtureFit:
    """One fitted weighted mixture: a device parameter block plus the scalars read back."""

    def __init__(self, block: torch.Tensor, d: int, k: int, offsets: np.ndarray):
        self.block, self.d, self.k, self.off = block, d, k, offsets
        self.n_iter = 0
        self.mean_bound = float("nan")     # sum_i (1/n) log(sum_k w_k N_k + 1e-10)
        self.n = 0

    def bic(self) -> float:                # cluster.py:310-340 (full covariances)
        d, k, n = self.d, self.k, self.n
        n_par = (k - 1) + k * d + k * d * (d + 1) / 2
        return -2 * (self.mean_bound * n) + n_par * math.log(n)

    def host_params(self):
        h = self.block.cpu().numpy()
        o, d, k = self.off, self.d, self.k
        return (h[o[0]:o[0] + k].copy(), h[o[1]:o[1] + k * d].reshape(k, d).copy(),
                h[o[2]:o[2] + k * d * d].reshape(k, d, d).copy())


class _Cluster:
    def __init__(self, members: Optional[torch.Tensor], size: int):
        self.members, self.size = members, size
        self.sw: Optional[torch.Tensor] = None     # normalised member weights
        self.weight_sum = float("nan")
        self.ess = float("nan")
        self.parent: Optional[MixtureFit] = None
        self.child: Optional[MixtureFit] = None
        self.gain = float("nan")
        self.split = None                           # (_Cluster, _Cluster) once computed
        self.evaluated = False


class HierarchicalGaussianMixture:
    """cluster.py:343-696 with ``covariance_type='full'`` (what core.py:59-69 constructs)."""

    def __init__(self, kernels, n_init=1, max_iterations=1000, min_points=None, threshold_modifier=1.0,
                 covariance_type="full", verbose=False, normalize=False):
        if covariance_type != "full":
            raise NotImplementedError("only covariance_type='full' (the sampler's setting) is built")
        if n_init != 1:
            raise NotImplementedError("n_init != 1 is not built (the sampler uses 1)")
        modifier = float(threshold_modifier)
        if modifier <= 0:
            raise ValueError("threshold_modifier must be positive.")      # cluster.py:361-362
        self.k = kernels
        self.lib = kernels.lib
        self.max_iterations = max_iterations
        self.min_points = min_points
        self.threshold_modifier = modifier
        self.verbose = verbose
        self.normalize = normalize
        self.labels_ = None
        self.cluster_centers_: List[np.ndarray] = []
        self.cluster_covariances_: List[np.ndarray] = []
        self.cluster_weights_ = None
        self.n_clusters_ = 0
        self._gmm_ready = False
        self._data_min = None
        self._data_max = None
        self._lo = self._hi = None
        self._predict_block = None
        self.n_fits = 0
        self.diagnostics: List[dict] = []

    # ---- one weighted EM fit (cluster.py:56-133) ------------------------------------------------
    def _fit_mixture(self, xn: torch.Tensor, cl: _Cluster, k: int, d: int) -> MixtureFit:
        lib, ws = self.lib, self.k.ws
        st = stream_ptr()
        n = cl.size
        rows = cl.members
        off = np.zeros(8, dtype=np.int64)
        _lib.check(lib.tb_gmm_offsets(d, k, off.ctypes.data), "tb_gmm_offsets")
        block = torch.zeros(int(off[7]), dtype=F64, device=xn.device)
        fit = MixtureFit(block, d, k, off)
        fit.n = n
        centres = ws.f64("gmm_centres", k * d)
        fracs = _seed_uniforms(k)
        # weighted k-means++ (cluster.py:139-158)
        run = self.k.cdf(cl.sw, n, "gmm_run")
        _lib.check(lib.tb_kpp_pick(ptr(run), n, float(fracs[0]), ptr(xn), ptr(rows), d, ptr(centres), None, st),
                   "tb_kpp_pick")
        for c in range(1, k):
            p = ws.f64("gmm_p", n)
            _lib.check(lib.tb_kpp_prob(ptr(xn), ptr(rows), ptr(cl.sw), n, d, ptr(centres), c, ptr(p), st), "tb_kpp_prob")
            stats = ws.f64("gmm_pstats", 3)
            _lib.check(lib.tb_normalize_inplace(ptr(p), n, ptr(self.k._reduce_ws), ptr(stats), st),
                       "tb_normalize_inplace")
            run = self.k.cdf(p, n, "gmm_run")
            _lib.check(lib.tb_kpp_pick(ptr(run), n, float(fracs[c]), ptr(xn), ptr(rows), d, ptr(centres[c * d:]),
                                       None, st), "tb_kpp_pick")
        wr = ws.f64("gmm_wr", k * n)
        gws = ws.bytes("gmm_ws", lib.tb_gmm_workspace_bytes())
        mws = ws.bytes("mom", lib.tb_moments_workspace_bytes(d))
        _lib.check(lib.tb_gmm_init(ptr(xn), ptr(rows), ptr(cl.sw), n, d, k, ptr(centres), ptr(block), ptr(wr), ptr(gws),
                                   ptr(mws), REG_COVAR, st), "tb_gmm_init")
        chunk = FIRST_CHUNK
        while True:
            _lib.check(lib.tb_gmm_em(ptr(xn), ptr(rows), ptr(cl.sw), n, d, k, ptr(block), ptr(wr), ptr(gws), ptr(mws),
                                     REG_COVAR, EM_TOL, EM_MAX_ITER, chunk, st), "tb_gmm_em")
            _lib.check(lib.tb_gmm_bound(ptr(xn), ptr(rows), None, n, d, k, ptr(block), ptr(gws), st), "tb_gmm_bound")
            hdr = block[:8].cpu().numpy()
            if hdr[1] != 0.0 or hdr[2] >= EM_MAX_ITER:
                break
            chunk = NEXT_CHUNK
        fit.n_iter = int(hdr[2])
        fit.mean_bound = float(hdr[7])
        self.n_fits += 1
        return fit

    def _weights_of(self, cl: _Cluster, sw_all: torch.Tensor) -> None:
        """Member weights, normalised (cluster.py:88), their sum and ESS (cluster.py:407-411)."""
        if cl.sw is not None:
            return
        lib = self.lib
        st = stream_ptr()
        sw = torch.empty(cl.size, dtype=F64, device=sw_all.device)
        if cl.members is None:
            sw.copy_(sw_all[:cl.size])
        else:
            _lib.check(lib.tb_take(ptr(sw_all), ptr(cl.members), cl.size, ptr(sw), st), "tb_take")
        stats = self.k.ws.f64("gmm_swstats", 3)
        _lib.check(lib.tb_normalize_inplace(ptr(sw), cl.size, ptr(self.k._reduce_ws), ptr(stats), st),
                   "tb_normalize_inplace")
        h = stats.cpu().numpy()
        cl.sw = sw
        cl.weight_sum = float(h[0])
        cl.ess = 1.0 / float(h[1])

    def _evaluate(self, xn, cl: _Cluster, sw_all, d: int) -> None:
        """Parent / child fits and the BIC gain of one cluster (cluster.py:456-479).  Results depend
        only on the members, so a cluster that survives a round is not refitted."""
        if cl.evaluated:
            return
        self._weights_of(cl, sw_all)
        cl.parent = self._fit_mixture(xn, cl, 1, d)
        cl.child = self._fit_mixture(xn, cl, 2, d)
        cl.gain = cl.parent.bic() - cl.child.bic()
        cl.evaluated = True

    def _split(self, xn, cl: _Cluster, d: int):
        """Children of the 2-component fit, member order preserved (cluster.py:494-496)."""
        if cl.split is None:
            lib = self.lib
            st = stream_ptr()
            n = cl.size
            labels = self.k.ws.i32("gmm_labels", n)
            _lib.check(lib.tb_gmm_predict(ptr(xn), ptr(cl.members), n, d, 2, ptr(cl.child.block), None, None, 1,
                                          ptr(labels), st), "tb_gmm_predict")
            zero = torch.empty(n, dtype=torch.int64, device=xn.device)
            one = torch.empty(n, dtype=torch.int64, device=xn.device)
            n_one = self.k.ws.i64("gmm_n_one", 1)
            sws = self.k.ws.bytes("gmm_split", lib.tb_split_workspace_bytes(n))
            _lib.check(lib.tb_split_by_label(ptr(labels), ptr(cl.members), n, ptr(sws), ptr(zero), ptr(one), ptr(n_one),
                                             st), "tb_split_by_label")
            c1 = int(n_one.item())
            c0 = n - c1
            cl.split = (_Cluster(zero[:c0], c0), _Cluster(one[:c1], c1))
        return cl.split

    # ---- cluster.py:420-572 ------------------------------------------------------------------------
    def fit(self, X: torch.Tensor, sample_weight: torch.Tensor, rows: Optional[torch.Tensor] = None):
        """``X``: row-major [., d] device tensor; ``rows`` (optional int64) selects and orders the
        samples (the trimmed set); ``sample_weight``: one weight per selected sample."""
        lib = self.lib
        st = stream_ptr()
        d = int(X.shape[1])
        n = int(rows.numel()) if rows is not None else int(X.shape[0])
        if int(sample_weight.numel()) != n:
            raise ValueError("sample_weight must have the same length as X")
        dev = X.device
        xn = torch.empty((n, d), dtype=F64, device=dev)
        if self.normalize:                                   # cluster.py:436-439
            self._lo = torch.empty(d, dtype=F64, device=dev)
            self._hi = torch.empty(d, dtype=F64, device=dev)
            mm = self.k.ws.bytes("gmm_minmax", lib.tb_col_minmax_workspace_bytes(d))
            _lib.check(lib.tb_col_minmax(ptr(X), ptr(rows), n, d, ptr(mm), ptr(self._lo), ptr(self._hi), st),
                       "tb_col_minmax")
            self._data_min = self._lo.cpu().numpy()
            self._data_max = self._hi.cpu().numpy()
            _lib.check(lib.tb_gather_normalised(ptr(X), ptr(rows), n, d, ptr(self._lo), ptr(self._hi), ptr(xn), st),
                       "tb_gather_normalised")
        else:
            self._lo = self._hi = None
            _lib.check(lib.tb_gather_normalised(ptr(X), ptr(rows), n, d, None, None, ptr(xn), st),
                       "tb_gather_normalised")
        need = self.min_points if self.min_points is not None else 2 * d       # :441
        clusters: List[_Cluster] = [_Cluster(None, n)]
        self.diagnostics = []
        rounds = 0
        while rounds < self.max_iterations:                  # :445
            rounds += 1
            best_gain, best_c, best_split = -math.inf, None, None
            diag = []
            for c, cl in enumerate(clusters):
                if cl.size < need:                           # :453
                    diag.append(None)
                    continue
                self._evaluate(xn, cl, sample_weight, d)
                n_par = d + d * (d + 1) / 2 + 1              # :413-418
                threshold = self.threshold_modifier * (n_par * math.log(cl.ess))
                diag.append(dict(size=cl.size, gain=cl.gain, threshold=threshold,
                                 parent_iter=cl.parent.n_iter, child_iter=cl.child.n_iter))
                if cl.gain > threshold and cl.gain > best_gain:                 # :493
                    a, b = self._split(xn, cl, d)
                    if a.size >= need and b.size >= need:    # :497
                        best_gain, best_c, best_split = cl.gain, c, (a, b)
            self.diagnostics.append(dict(clusters=diag, split=best_c))
            if best_split is None:
                break
            clusters.pop(best_c)                             # :508-509
            clusters.extend(best_split)
        # final per-cluster centre / covariance (:523-556)
        labels = torch.full((n,), -1, dtype=torch.int32, device=dev)
        centres, covs, sums = [], [], []
        for c, cl in enumerate(clusters):
            self._weights_of(cl, sample_weight)
            sums.append(cl.weight_sum)
            if cl.size >= d:                                 # :531
                if cl.parent is None:
                    cl.parent = self._fit_mixture(xn, cl, 1, d)
                _, mean, cov = cl.parent.host_params()
                centre, cov = mean[0], cov[0]
            else:                                            # :545-547
                rows_c = cl.members if cl.members is not None else torch.arange(n, device=dev)
                centre = xn[rows_c].mean(dim=0).cpu().numpy()
                cov = np.eye(d)
            if self.normalize:                               # :385-405
                scale = self._data_max - self._data_min
                centre = centre * scale + self._data_min
                cov = cov * np.outer(scale, scale)
            centres.append(centre)
            covs.append(cov)
            if cl.members is None:
                labels.fill_(c)
            else:
                labels[cl.members] = c
        self.labels_ = labels
        self.cluster_centers_ = centres
        self.cluster_covariances_ = covs
        self.n_clusters_ = len(clusters)
        total = float(np.sum(np.array(sums)))                # :563-569
        self.cluster_weights_ = np.array(sums) / total
        self._gmm_ready = self.n_clusters_ > 0
        self._build_predictor(d, dev)
        return self

    def _build_predictor(self, d: int, dev) -> None:
        """Parameter block for ``_compute_gaussian_probabilities`` (cluster.py:633-691): normalised
        means and covariances, ``+ 1e-6 I``, identity fallback."""
        k = self.n_clusters_
        off = np.zeros(8, dtype=np.int64)
        _lib.check(self.lib.tb_gmm_offsets(d, k, off.ctypes.data), "tb_gmm_offsets")
        h = np.zeros(int(off[7]))
        for c in range(k):
            if self.normalize:                               # :645-651
                scale = self._data_max - self._data_min
                mean = (self.cluster_centers_[c] - self._data_min) / (scale + 1e-10)
                cov = self.cluster_covariances_[c] / np.outer(scale, scale)
            else:
                mean, cov = self.cluster_centers_[c], self.cluster_covariances_[c]
            h[off[0] + c] = self.cluster_weights_[c]
            h[off[1] + c * d: off[1] + (c + 1) * d] = mean
            h[off[2] + c * d * d: off[2] + (c + 1) * d * d] = np.asarray(cov).reshape(-1)
        block = torch.from_numpy(h).to(dev)
        _lib.check(self.lib.tb_gmm_prepare(ptr(block), d, k, 1e-6, 1.0, stream_ptr()), "tb_gmm_prepare")
        self._predict_block = block
        self._predict_d = d

    # -- checkpointing: everything predict() needs (the fit itself is not resumable mid-way, nor does it need to be)
    def state_dict(self) -> dict:
        if not self._gmm_ready:
            return {}
        out = dict(clu_centres=np.array(self.cluster_centers_), clu_covs=np.array(self.cluster_covariances_),
                   clu_weights=np.array(self.cluster_weights_), clu_d=np.array(self._predict_d))
        if self._data_min is not None:
            out.update(clu_min=np.array(self._data_min), clu_max=np.array(self._data_max))
        return out

    def load_state_dict(self, d: dict, device) -> None:
        if "clu_centres" not in d:
            return
        self.cluster_centers_ = [np.array(c) for c in d["clu_centres"]]
        self.cluster_covariances_ = [np.array(c) for c in d["clu_covs"]]
        self.cluster_weights_ = np.array(d["clu_weights"])
        self.n_clusters_ = len(self.cluster_centers_)
        dim = int(d["clu_d"])
        if "clu_min" in d:
            self._data_min, self._data_max = np.array(d["clu_min"]), np.array(d["clu_max"])
            self._lo = torch.as_tensor(self._data_min, dtype=F64).to(device)
            self._hi = torch.as_tensor(self._data_max, dtype=F64).to(device)
        else:
            self._data_min = self._data_max = None
            self._lo = self._hi = None
        self._gmm_ready = self.n_clusters_ > 0
        self._build_predictor(dim, device)

    def predict(self, X: torch.Tensor, rows: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None):
        """cluster.py:574-600: int32 device labels of ``X[rows]``."""
        if not self._gmm_ready:
            raise ValueError("The model has not been fitted yet.")
        d = self._predict_d
        n = int(rows.numel()) if rows is not None else int(X.shape[0])
        if out is None:
            out = torch.empty(n, dtype=torch.int32, device=X.device)
        _lib.check(self.lib.tb_gmm_predict(ptr(X), ptr(rows), n, d, self.n_clusters_, ptr(self._predict_block),
                                           ptr(self._lo), ptr(self._hi), 0, ptr(out), stream_ptr()),
                   "tb_gmm_predict")
        return out
