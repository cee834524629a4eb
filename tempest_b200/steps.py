"""Device implementations of the four Persistent Sampling steps.

Mirrors ``Reweighter / Trainer / Resampler / Mutator .run()`` of the reference
(tempest/steps/{reweight,train,resample,mutate}.py) -- same state keys written, same
constants, same branch logic -- with every array operation executed by libtempest_b200
kernels through the C ABI.  Host code only sequences kernels and takes the scalar decisions
the reference takes in Python.
"""
from __future__ import annotations

import ctypes as C
import math
import weakref
from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .config import (BETA_RTOL, BETA_TOLERANCE, DOF_FALLBACK, ESS_TOLERANCE, MAX_BISECTION_ITERATIONS,
                     METRIC_ATOL, METRIC_ATOL_CV, TRIM_BINS, TRIM_ESS)
from .ensemble import PersistentEnsemble, ptr, stream_ptr

_TINY = float(np.finfo(float).tiny)
_EPS = float(np.finfo(float).eps)
F64 = torch.float64


class Workspace:
    """Reusable device scratch (the C ABI never allocates)."""

    def __init__(self, device: torch.device):
        self.device = device
        self.lib = _lib.load()
        self._buf = {}
        self.hint = 0          # capacity of the persistent ensemble: N_total-sized scratch is allocated once

    def _size(self, n: int) -> int:
        if self.hint and n > self.hint // 64:
            return max(int(n), self.hint) if n <= self.hint else int(n * 2)
        return max(int(n * 2), 16)

    def bytes(self, name: str, nbytes: int) -> torch.Tensor:
        t = self._buf.get(name)
        if t is None or t.numel() < nbytes:
            t = torch.zeros(max(int(nbytes) * 2, 256), dtype=torch.uint8, device=self.device)
            self._buf[name] = t
        return t

    def f64(self, name: str, n: int) -> torch.Tensor:
        t = self._buf.get(name)
        if t is None or t.numel() < n:
            t = torch.empty(self._size(n), dtype=F64, device=self.device)
            self._buf[name] = t
        return t[:n]

    def i64(self, name: str, n: int) -> torch.Tensor:
        t = self._buf.get(name)
        if t is None or t.numel() < n:
            t = torch.empty(self._size(n), dtype=torch.int64, device=self.device)
            self._buf[name] = t
        return t[:n]

    def i32(self, name: str, n: int) -> torch.Tensor:
        t = self._buf.get(name)
        if t is None or t.numel() < n:
            t = torch.empty(self._size(n), dtype=torch.int32, device=self.device)
            self._buf[name] = t
        return t[:n]


# ======================================================================================
# numpy arithmetic the reference relies on, restated for the host-side scalar decisions
# ======================================================================================
def numpy_pairwise_sum_equal(value: float, n: int) -> float:
    """``np.sum(np.full(n, value))`` without materialising the array: numpy's pairwise
    summation (8 running lanes over blocks of <= 128, halves split at ``n/2 - (n/2) % 8``)."""
    memo = {}

    def rec(m: int) -> float:
        if m in memo:
            return memo[m]
        if m < 8:
            r = 0.0
            for _ in range(m):
                r += value
        elif m <= 128:
            lanes = [value] * 8
            i = 8
            while i < m - (m % 8):
                for j in range(8):
                    lanes[j] += value
                i += 8
            r = ((lanes[0] + lanes[1]) + (lanes[2] + lanes[3])) + ((lanes[4] + lanes[5]) + (lanes[6] + lanes[7]))
            while i < m:
                r += value
                i += 1
        else:
            m2 = m // 2
            m2 -= m2 % 8
            r = rec(m2) + rec(m - m2)
        memo[m] = r
        return r

    return rec(int(n))


def uniform_weights_ess(n: int) -> float:
    """ESS the reference computes for ``n`` exactly equal weights (warm-up):
    ``w = ones(n)``; ``w / np.sum(w)``; ``1 / np.sum(w**2)`` (tools.py:134-135).  Equals ``n``
    only when the roundings cancel -- e.g. n = 42 gives 42.000000000000007 and the reference
    then leaves the warm-up one generation early, so the branch must see the same value."""
    wn = 1.0 / float(n)        # np.sum(ones(n)) is exact
    return 1.0 / numpy_pairwise_sum_equal(wn * wn, n)


def percentile_position(n: int, p: float) -> Tuple[int, int, float]:
    """numpy ``percentile(..., method='linear')`` virtual index -> (lo, hi, gamma)."""
    v = (n - 1) * (p / 100.0)
    lo = int(math.floor(v))
    g = v - lo
    hi = min(lo + 1, n - 1)
    lo = min(max(lo, 0), n - 1)
    return lo, hi, g


def percentile_positions(n: int, percentiles: np.ndarray):
    """Vector form of :func:`percentile_position` (same IEEE operations, element by element)."""
    v = (n - 1) * (percentiles / 100.0)
    lo = np.floor(v)
    g = v - lo
    lo = lo.astype(np.int64)
    hi = np.minimum(lo + 1, n - 1)
    lo = np.minimum(np.maximum(lo, 0), n - 1)
    return lo, hi, g


def numpy_lerp(a: float, b: float, t: float) -> float:
    """numpy ``_lerp``: ``a + (b-a) t`` switched to ``b - (b-a)(1-t)`` for t >= 0.5."""
    diff = b - a
    if t >= 0.5:
        return b - diff * (1 - t)
    return a + diff * t


# ======================================================================================
# shared device helpers
# ======================================================================================
# Workspaces are pooled per process and device: a Sampler leases one and hands it back when it is garbage-collected, so the
# next Sampler finds every named scratch buffer already allocated.  On a multi-GPU box a cudaMalloc has to map the new
# block into every peer (5-10 ms each with peer access enabled): the four or five a fresh workspace needs in its first
# trained iteration were a 30 ms spike at 8 GPUs.  Kernels never rely on workspace contents surviving between calls
# (tickets / counters reset themselves, status words are cleared per call), so stale contents are harmless.
_WS_POOL: dict = {}


def _ws_acquire(device: torch.device, role: str) -> "Workspace":
    # keyed by role: the main-stream and the side-stream workspace of a Sampler hold different named buffers, and handing
    # one out as the other would re-allocate everything while the pool still holds the blocks
    idle = _WS_POOL.setdefault((device.index, role), [])
    ws = idle.pop() if idle else Workspace(device)
    ws.hint = 0
    return ws


def _ws_release(key, ws) -> None:
    _WS_POOL.setdefault(key, []).append(ws)


class Kernels:
    """Thin typed wrappers over the C ABI bound to one device / workspace."""

    def __init__(self, device: torch.device, role: str = "main"):
        self.device = device
        self.lib = _lib.load()
        self.ws = _ws_acquire(device, role)
        weakref.finalize(self, _ws_release, (device.index, role), self.ws)
        self._probe_ws = self.ws.bytes("probe", self.lib.tb_probe_workspace_bytes())
        self._reduce_ws = self.ws.bytes("reduce", self.lib.tb_reduce_workspace_bytes())
        self.probe_out = torch.zeros(16, dtype=F64, device=device)
        self.n_probe_launches = 0
        self.sharded = False

    # -- reweighting ----------------------------------------------------------------------
    def probe(self, ens: PersistentEnsemble, beta: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        out = self.probe_out if out is None else out
        _lib.check(self.lib.tb_probe(ptr(ens.logl), ptr(ens.C), ens.n_total, float(beta), ptr(self._probe_ws),
                                     ptr(out), stream_ptr()), "tb_probe")
        self.n_probe_launches += 1
        return out

    def next_beta(self, ens: PersistentEnsemble, beta_prev: float, target: float, flags: int,
                  log_cap: int = 512):
        res = self.ws.f64("nb_res", 16)
        plog = self.ws.f64("nb_log", 2 * log_cap)
        _lib.check(self.lib.tb_next_beta(ptr(ens.logl), ptr(ens.C), ens.n_total, float(beta_prev), float(target),
                                         int(flags), ptr(self._probe_ws), ptr(res), ptr(plog), log_cap,
                                         stream_ptr()), "tb_next_beta")
        return res, plog

    def weights(self, ens: PersistentEnsemble, beta: float, stats: torch.Tensor, out: torch.Tensor, log=False):
        fn = self.lib.tb_log_weights if log else self.lib.tb_weights
        _lib.check(fn(ptr(ens.logl), ptr(ens.C), ens.n_total, float(beta), ptr(stats), ptr(out), stream_ptr()),
                   "tb_weights")
        return out

    # -- moments / cv -------------------------------------------------------------------
    def volume_variation(self, u: torch.Tensor, w: torch.Tensor, n: int, d: int) -> float:
        """tools.py:58-117 on the device (w already normalised)."""
        self.volume_variation_begin(u, w, n, d)
        self.volume_variation_mid(u, w, n, d)
        return self.volume_variation_end()

    # The three phases can be issued on a side stream around other host work: `begin` only enqueues
    # (weighted moments + Cholesky), `mid` takes the rank decision on the host and enqueues the
    # Mahalanobis pass, `end` reads the scalar.
    def volume_variation_begin(self, u: torch.Tensor, w: torch.Tensor, n: int, d: int) -> None:
        self._vv_state = "small" if n < d + 1 else "begun"
        if n < d + 1:
            return
        ws = self.ws.bytes("mom", self.lib.tb_moments_workspace_bytes(d))
        mean = self.ws.f64("vv_mean", d)
        cov = self.ws.f64("vv_cov", d * d)
        _lib.check(self.lib.tb_weighted_moments(ptr(u), ptr(w), n, d, ptr(ws), ptr(mean), ptr(cov), stream_ptr()),
                   "tb_weighted_moments")
        work = self.ws.f64("vv_work", d * d)
        inv = self.ws.f64("vv_inv", d * d)
        info = self.ws.i32("vv_info", 1)
        norms = self.ws.f64("vv_norms", 3)
        work.copy_(cov)
        _lib.check(self.lib.tb_chol_inv(ptr(work), d, 1, None, ptr(inv), ptr(info), ptr(norms), stream_ptr()),
                   "tb_chol_inv")

    def volume_variation_mid(self, u: torch.Tensor, w: torch.Tensor, n: int, d: int) -> None:
        if self._vv_state != "begun":
            return
        mean, cov = self.ws.f64("vv_mean", d), self.ws.f64("vv_cov", d * d)
        work, inv = self.ws.f64("vv_work", d * d), self.ws.f64("vv_inv", d * d)
        info, norms = self.ws.i32("vv_info", 1), self.ws.f64("vv_norms", 3)
        for attempt in range(2):
            if attempt == 1:
                work.copy_(cov)
                _lib.check(self.lib.tb_chol_inv(ptr(work), d, 1, None, ptr(inv), ptr(info), ptr(norms), stream_ptr()),
                           "tb_chol_inv")
            code = int(info.item())
            nr = norms.cpu().numpy()
            # matrix_rank(cov) < d  <=>  cond_2 >= 1/(d*eps); |A|_F |A^-1|_F >= cond_2 screens it
            if attempt == 0:
                singular = code != 0 or not np.isfinite(nr[:2]).all() or nr[0] * nr[1] >= 1.0 / (d * _EPS)
            else:
                singular = code == 2 or not np.isfinite(nr[:2]).all()   # np.linalg.inv raised (tools.py:108-110)
            if not singular:
                break
            if attempt == 1:
                self._vv_state = "singular"
                return
            _lib.check(self.lib.tb_add_trace_reg(ptr(cov), d, 1e-6, stream_ptr()), "tb_add_trace_reg")
        out = self.ws.f64("vv_out", 2)
        _lib.check(self.lib.tb_mahalanobis_cv(ptr(u), ptr(w), n, d, ptr(mean), ptr(inv), ptr(self._reduce_ws),
                                              ptr(out), stream_ptr()), "tb_mahalanobis_cv")
        self._vv_state = "enqueued"

    def volume_variation_end(self) -> float:
        if self._vv_state in ("small", "singular"):
            return 1e10
        return float(self.ws.f64("vv_out", 2)[0].item())

    # -- resampling -----------------------------------------------------------------------
    def cdf(self, p: torch.Tensor, n: int, name: str = "cdf") -> torch.Tensor:
        out = self.ws.f64(name, n)
        ws = self.ws.bytes("cdf_ws", self.lib.tb_cdf_workspace_bytes(n))
        _lib.check(self.lib.tb_cdf_exact(ptr(p), n, ptr(out), ptr(ws), stream_ptr()), "tb_cdf_exact")
        return out

    def search_right(self, cdf: torch.Tensor, n: int, draws: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        m = int(draws.numel())
        if m >= 1 << 16 and n >= 1 << 12:
            # many draws: bracket each one with a guide table (identical indices, ~3 probes instead of log2 n)
            bits = max(10, min(20, (m // 8).bit_length() - 1))
            guide = self.ws.bytes("search_guide", self.lib.tb_search_guide_bytes(bits))
            _lib.check(self.lib.tb_search_right_guided(ptr(cdf), n, ptr(draws), m, ptr(guide), bits, ptr(out),
                                                       stream_ptr()), "tb_search_right_guided")
            return out
        _lib.check(self.lib.tb_search_right(ptr(cdf), n, ptr(draws), m, ptr(out), stream_ptr()), "tb_search_right")
        return out

    def systematic(self, cdf: torch.Tensor, n: int, u0: float, m: int, out: torch.Tensor) -> torch.Tensor:
        flag = self.ws.i32("syst_flag", 1)
        _lib.check(self.lib.tb_systematic(ptr(cdf), n, float(u0), m, ptr(out), ptr(flag), stream_ptr()),
                   "tb_systematic")
        if int(flag.item()):
            raise IndexError("systematic resampling walked past the last weight (tools.py:223-225)")
        return out

    # -- global reductions (overridden by sharded.ShardedKernels with all-reduces) -------------------
    def g_int(self, value: int) -> int:
        return int(value)

    def g_normalize(self, w: torch.Tensor, n: int):
        """w /= sum(w) in place (tools.py:36); returns (sum before, sum of squares after)."""
        return self.g_normalize_end(self.g_normalize_begin(w, n))

    # split form: `begin` only enqueues (the sum stays on the device), `end` reads the two scalars -- callers that do
    # not need them skip `end`, callers that do read them together with their next host read
    def g_normalize_begin(self, w: torch.Tensor, n: int):
        stats = self.ws.f64("trim_stats", 3)
        _lib.check(self.lib.tb_normalize_inplace(ptr(w), n, ptr(self._reduce_ws), ptr(stats), stream_ptr()),
                   "tb_normalize_inplace")
        return stats

    def g_normalize_end(self, stats):
        h = stats.cpu().numpy()
        return float(h[0]), float(h[1])

    def _hist_buffers(self, name: str):
        """count / sum w / sum w^2 for 2048 bins in ONE allocation, so one copy brings all three to the host."""
        buf = self.ws.f64(name, 3 * 2048)
        return buf, buf[:2048].view(torch.int64), buf[2048:4096], buf[4096:]

    @staticmethod
    def _hist_to_host(buf: torch.Tensor):
        h = buf.cpu().numpy()
        return h[:2048].view(np.int64).copy(), h[2048:4096], h[4096:]

    def g_hist(self, w: torch.Tensor, n: int):
        buf, cnt, s1, s2 = self._hist_buffers("trim_hist")
        _lib.check(self.lib.tb_binade_hist(ptr(w), n, ptr(cnt), ptr(s1), ptr(s2), stream_ptr()), "tb_binade_hist")
        return self._hist_to_host(buf)

    def g_subhist(self, w: torch.Tensor, n: int, binade: int):
        buf, cnt, s1, s2 = self._hist_buffers("trim_hist2")
        _lib.check(self.lib.tb_subbin_hist(ptr(w), n, int(binade), ptr(cnt), ptr(s1), ptr(s2), stream_ptr()),
                   "tb_subbin_hist")
        return self._hist_to_host(buf)

    def g_sum3(self, vals: torch.Tensor, m: int, thr: float):
        """(count, sum w, sum w^2) over w >= thr."""
        out3 = self.ws.f64("trim_m3", 3)
        if m > 0:
            _lib.check(self.lib.tb_masked_sums(ptr(vals), m, float(thr), ptr(self._reduce_ws), ptr(out3),
                                               stream_ptr()), "tb_masked_sums")
        else:
            out3.zero_()
        c, a1, a2 = out3.cpu().numpy()
        return float(c), float(a1), float(a2)

    def g_select(self, base, rows, stride: int, m: int, ncols: int, mult, ranks: torch.Tensor, nranks: int,
                 out: torch.Tensor) -> torch.Tensor:
        sws = self.ws.bytes("select", self.lib.tb_select_workspace_bytes(ncols, nranks))
        _lib.check(self.lib.tb_select_ranks(ptr(base), ptr(rows), stride, m, ncols, ptr(mult), ptr(ranks), nranks,
                                            ptr(sws), ptr(out), stream_ptr()), "tb_select_ranks")
        return out

    def g_select_pair(self, base, rows, stride: int, m: int, ncols: int, mult, rank_lo: int, same: bool,
                      out: torch.Tensor) -> torch.Tensor:
        """Order statistics (rank_lo, rank_lo + 1) of every column -> out[2c], out[2c+1]."""
        sws = self.ws.bytes("select_pair", self.lib.tb_select_pair_workspace_bytes(ncols))
        _lib.check(self.lib.tb_select_pair(ptr(base), ptr(rows), stride, m, ncols, ptr(mult), int(rank_lo), int(same),
                                           ptr(sws), ptr(out), stream_ptr()), "tb_select_pair")
        return out

    # -- trim_weights (tools.py:10-55) ------------------------------------------------------
    def trim(self, w: torch.Tensor, n: int, ess: float = TRIM_ESS, bins: int = TRIM_BINS, after_normalize=None,
             n_global: Optional[int] = None):
        """Normalises ``w`` IN PLACE (tools.py:36) and returns (idx int64[n_trim], w_trim[n_trim]).

        The reference scans i = bins-1 .. 0, each time thresholding at np.percentile(w, p_i) and
        stopping at the first i whose trimmed ESS ratio reaches ``ess``.  The ratio is monotone in
        the threshold, so the same i is found by (1) one binade histogram pass that brackets the
        flip, (2) exact evaluations (radix order statistics + masked sums) of the few percentile
        grid points inside the bracket, by bisection.  ``n`` is the LOCAL length; every count,
        rank and sum below is global (the g_* hooks all-reduce when the ensemble is sharded)."""
        lib = self.lib
        st = stream_ptr()
        pending = self.g_normalize_begin(w, n)
        if after_normalize is not None:
            after_normalize()                 # w is final from here on (read-only below)
        h_cnt, h_s1, h_s2 = self.g_hist(w, n)
        _, sumsq = self.g_normalize_end(pending)          # (no second wait: the histogram read drained the stream)
        n_glob = int(n_global) if n_global is not None else self.g_int(n)
        ess_total = 1.0 / sumsq
        # suffix sums over binades (threshold at the lower edge of binade b keeps bins >= b)
        s1_ge = np.cumsum(h_s1[::-1])[::-1]
        s2_ge = np.cumsum(h_s2[::-1])[::-1]
        with np.errstate(divide="ignore", invalid="ignore"):
            ratio = (s1_ge * s1_ge / s2_ge) / ess_total
        ok_edge = np.nan_to_num(ratio, nan=0.0) >= ess
        bstar = int(np.max(np.nonzero(ok_edge)[0]))          # highest binade edge that still passes
        cnt_lt = np.concatenate([[0], np.cumsum(h_cnt)])      # elements in bins < b
        percentiles = np.linspace(0, 99, bins)
        lo_arr, hi_arr, g_arr = percentile_positions(n_glob, percentiles)      # a Python loop here cost 0.8 ms per call
        bin_lo = np.searchsorted(cnt_lt, lo_arr, side="right") - 1
        bin_hi = np.searchsorted(cnt_lt, hi_arr, side="right") - 1
        sure_true = bin_hi < bstar
        sure_false = bin_lo > bstar
        fine_edge = None
        if np.count_nonzero(~sure_true & ~sure_false) > 2 and 0 < bstar < 2047:
            # second level: 2048 linear sub-bins of binade bstar narrow the bracket to one sub-bin
            f_cnt, f_s1, f_s2 = self.g_subhist(w, n, bstar)
            a1 = (s1_ge[bstar + 1] if bstar + 1 < 2048 else 0.0) + np.cumsum(f_s1[::-1])[::-1]
            a2 = (s2_ge[bstar + 1] if bstar + 1 < 2048 else 0.0) + np.cumsum(f_s2[::-1])[::-1]
            with np.errstate(divide="ignore", invalid="ignore"):
                f_ok = np.nan_to_num((a1 * a1 / a2) / ess_total, nan=0.0) >= ess
            if f_ok.any() and int(f_cnt.sum()) == int(h_cnt[bstar]):
                kstar = int(np.max(np.nonzero(f_ok)[0]))         # highest sub-bin edge that still passes
                f_lt = cnt_lt[bstar] + np.concatenate([[0], np.cumsum(f_cnt)])   # elements below sub-edge k
                f_lo = np.searchsorted(f_lt, lo_arr, side="right") - 1
                f_hi = np.searchsorted(f_lt, hi_arr, side="right") - 1
                sure_true = sure_true | ((bin_hi == bstar) & (f_hi < kstar))
                sure_false = sure_false | ((bin_lo == bstar) & (f_lo > kstar))
                fine_edge = (f_lo, kstar)
        amb = np.nonzero(~sure_true & ~sure_false)[0]
        best = int(np.max(np.nonzero(sure_true)[0])) if sure_true.any() else 0
        cand = np.union1d(amb, [best]).astype(np.int64)   # grid points that may need an exact evaluation

        cache = {}
        comp = None  # (values tensor, local count, global count below the compacted set)
        hist_built = [False]
        comp_lo = 0.0
        nzb = np.nonzero(h_cnt)[0]
        w_max = float(np.ldexp(1.0, int(nzb.max()) - 1022)) if nzb.size else 1.0     # upper edge of the top binade

        def exact(i: int):
            nonlocal comp, comp_lo
            if i in cache:
                return cache[i]
            if comp is None:
                # compact everything from the lowest binade a candidate threshold can touch
                b0 = int(bin_lo[cand].min())
                if b0 <= 0:
                    comp = (w, n, 0)
                else:
                    edge = float(np.ldexp(1.0, b0 - 1023))
                    if fine_edge is not None and b0 == bstar:
                        # every candidate rank lies at or above this sub-bin edge of binade bstar
                        k0 = int(fine_edge[0][cand][bin_lo[cand] == bstar].min())
                        edge = float(np.ldexp(1.0 + max(k0, 0) / 2048.0, bstar - 1023))
                    comp_lo = edge
                    wc = self.ws.f64("trim_comp", n)
                    nout = self.ws.i64("trim_nout", 1)
                    cws = self.ws.bytes("compact", lib.tb_compact_workspace_bytes(n))
                    _lib.check(lib.tb_compact_ge(ptr(w), n, edge, 1.0, ptr(cws), None, ptr(wc), ptr(nout), st),
                               "tb_compact_ge")
                    m = int(nout.item())
                    comp = (wc, m, n_glob - self.g_int(m))
            vals, m, below = comp
            lo, hi, g = int(lo_arr[i]), int(hi_arr[i]), float(g_arr[i])
            sel = self.ws.f64("trim_sel", 2)
            done = False
            if not self.sharded and m > 0:
                # bucket histogram over the compacted tail is built once and reused by every evaluation
                bws = self.ws.bytes("bucket_pair", lib.tb_unit_median_workspace_bytes(1))
                ovf = self.ws.i32("bucket_ovf", 1)
                _lib.check(lib.tb_bucket_select_pair(ptr(vals), m, lo - below, int(hi == lo), comp_lo, w_max,
                                                     int(not hist_built[0]), ptr(bws), ptr(sel), ptr(ovf), st),
                           "tb_bucket_select_pair")
                hist_built[0] = True
                done = int(ovf.item()) == 0
            elif self.sharded:
                done = self.bucket_pair_sharded(vals if m > 0 else None, None, None, m, 1, lo - below, hi == lo, comp_lo,
                                                w_max, False, not hist_built[0], sel)
                hist_built[0] = True
            if not done:
                self.g_select_pair(vals, None, 1, m, 1, None, lo - below, hi == lo, sel)
            a, b = sel.cpu().numpy()
            thr = numpy_lerp(float(a), float(b), g)
            c, a1, a2 = self.g_sum3(vals, m, thr)
            good = ((a1 * a1 / a2) / ess_total) >= ess
            cache[i] = (bool(good), thr)
            return cache[i]

        chosen = best
        lo_i, hi_i = 0, amb.size - 1              # largest ambiguous grid point that passes (monotone)
        while lo_i <= hi_i:
            mid = (lo_i + hi_i) // 2
            if exact(int(amb[mid]))[0]:
                chosen = int(amb[mid])
                lo_i = mid + 1
            else:
                hi_i = mid - 1
        thr = exact(chosen)[1]
        # final: deterministic sum over the kept set on the full array, then ordered compaction
        c_glob, a1, _ = self.g_sum3(w, n, thr)
        nout = self.ws.i64("trim_nout", 1)
        cws = self.ws.bytes("compact", lib.tb_compact_workspace_bytes(n))
        if self.sharded:     # local kept count differs from the global one
            _lib.check(lib.tb_compact_ge(ptr(w), n, thr, float(a1), ptr(cws), None, None, ptr(nout), st),
                       "tb_compact_ge")
            n_loc = int(nout.item())
        else:
            n_loc = int(c_glob)
        idx = torch.empty(n_loc, dtype=torch.int64, device=self.device)
        wt = torch.empty(n_loc, dtype=F64, device=self.device)
        if n_loc:
            _lib.check(lib.tb_compact_ge(ptr(w), n, thr, float(a1), ptr(cws), ptr(idx), ptr(wt), ptr(nout), st),
                       "tb_compact_ge")
        self.last_trim = dict(bin=chosen, threshold=thr, n_trim=int(c_glob), n_exact=len(cache))
        return idx, wt


class ModeStats:
    """Device mirror of ``ModeStatistics`` (tempest/modes.py:58-119): per-mode mean, covariance,
    dof, lower Cholesky factor and inverse, all resident in HBM."""

    def __init__(self, means: torch.Tensor, covs: torch.Tensor, chol: torch.Tensor, inv: torch.Tensor,
                 dofs: torch.Tensor):
        self.means, self.covariances, self.chol_covariances, self.inv_covariances = means, covs, chol, inv
        self.degrees_of_freedom = dofs

    @property
    def K(self) -> int:
        return int(self.means.shape[0])

    @property
    def n_dim(self) -> int:
        return int(self.means.shape[1])

    @classmethod
    def identity(cls, d: int, device) -> "ModeStats":
        eye = torch.eye(d, dtype=F64, device=device).reshape(1, d, d)
        return cls(torch.zeros((1, d), dtype=F64, device=device), eye.clone(), eye.clone(), eye.clone(),
                   torch.full((1,), DOF_FALLBACK, dtype=F64, device=device))


# ======================================================================================
# The four steps
# ======================================================================================
class Reweighter:
    """steps/reweight.py:341-495."""

    def __init__(self, core):
        self.core = weakref.proxy(core)      # the core owns the steps; no reference cycle, so memory frees on del
        cfg = core.config
        self.n_particles = cfg.n_particles
        self.ess_ratio = cfg.ess_ratio
        self.volume_variation = cfg.volume_variation
        self.device_search = True     # tb_next_beta (one launch) vs host-driven probes
        self.SPECULATIVE_MAX = 4 << 20  # particles in this rank's history below which passes evaluate three betas
        self.probe_log: List[Tuple[float, float]] = []

    # one probe: returns (ess, metric) and leaves (m, S1, ...) in core.k.probe_out
    def _probe(self, beta: float) -> Tuple[float, float]:
        core = self.core
        ens = core.ensemble
        out = core.k.probe(ens, beta).cpu().numpy()
        ess = float(out[3])
        if core.warmup_regime and beta == 0.0:
            ess = uniform_weights_ess(ens.n_total_global)
        metric = ess
        if self.volume_variation is not None:
            w = core.k.weights(ens, beta, core.k.probe_out, core.weights_buffer())
            metric = core.k.volume_variation(ens.u, w, ens.n_total, ens.n_dim)
        self.probe_log.append((float(beta), ess))
        return ess, metric

    def _ess_bracket(self, beta_current: float, target: float) -> Tuple[float, float]:
        lo, hi = beta_current, 1.0
        ess_cur, _ = self._probe(beta_current)
        if ess_cur <= target:                      # reweight.py:264-266
            return beta_current, beta_current
        ess_one, _ = self._probe(1.0)
        if ess_one >= target:                      # :269-271
            return 1.0, 1.0
        while True:                                # :277-295
            mid = (hi + lo) * 0.5
            scale = max(abs(lo), abs(hi), _TINY)
            if hi - lo <= max(BETA_RTOL * scale, BETA_TOLERANCE * scale):
                break
            ess_mid, _ = self._probe(mid)
            if ess_mid >= target:
                lo = mid
            else:
                hi = mid
        return lo, hi

    def _bisect(self, beta_min: float, beta_max: float, target: float, use_metric: bool):
        dynamic = self.volume_variation is not None
        beta = beta_min
        ess = float("nan")
        for _ in range(MAX_BISECTION_ITERATIONS):  # reweight.py:162-223
            beta = (beta_max + beta_min) * 0.5
            ess, metric = self._probe(beta)
            val = metric if use_metric else ess
            if not math.isfinite(val):
                val = 1e10
            atol = METRIC_ATOL_CV if dynamic else METRIC_ATOL
            metric_ok = abs(val - target) < max(ESS_TOLERANCE * abs(target), atol)
            scale = max(abs(beta_min), abs(beta_max), _TINY)
            beta_ok = (beta_max - beta_min) < max(BETA_RTOL * scale, BETA_TOLERANCE * scale)
            if metric_ok or beta_ok or beta == 1.0:
                return beta, ess
            if not use_metric:
                if val < target:
                    beta_max = beta
                else:
                    beta_min = beta
            else:
                if val < target:
                    beta_min = beta
                else:
                    beta_max = beta
        return beta, ess

    def run(self) -> Optional[torch.Tensor]:
        core = self.core
        st = core.state
        ens = core.ensemble
        st.set_current("iter", st.raw("iter") + 1)
        self.probe_log = []
        self._logz_host = None
        n = self.n_particles
        if ens.T == 0:                              # reweight.py:365-383
            st.update_current({"beta": 0.0, "logz": 0.0, "ess": self.ess_ratio * n, "cv": 0.0})
            return None                             # uniform 1/N weights, never consumed at beta = 0
        beta_prev = float(st.raw("beta"))
        target = self.ess_ratio * n
        dynamic = self.volume_variation is not None
        k = core.k
        if not dynamic and self.device_search and (not k.sharded or k.xgpu is not None):
            beta, ess, stats = self._device_search(beta_prev, target)
        else:
            lo, hi = self._ess_bracket(beta_prev, target)
            if lo == hi:
                beta = lo
                ess, _ = self._probe(beta)
            elif not dynamic:
                beta, ess = self._bisect(beta_prev, hi, target, use_metric=False)
            else:                                   # reweight.py:427-482
                ess_prev, cv_prev = self._probe(beta_prev)
                ess_high, cv_high = self._probe(hi)
                need_probe = True
                if self.volume_variation >= cv_high:
                    beta, ess = hi, ess_high
                elif self.volume_variation <= cv_prev:
                    beta, ess = beta_prev, ess_prev
                else:
                    beta, ess = self._bisect(beta_prev, hi, self.volume_variation, use_metric=True)
                    need_probe = False
                if need_probe:
                    ess, _ = self._probe(beta)
            stats = k.probe_out                     # (m, S1, ..., logZ) of the last probe == probe(beta)
        core._stage("reweight:cv")
        w = k.weights(ens, beta, stats, core.weights_buffer())
        if core.overlap:
            core.begin_cv(w)                                          # side stream; finished before the commit
            cv = None
        elif k.sharded and not dynamic:
            k.volume_variation_async(ens.u, w, ens.n_total, ens.n_dim)    # enqueued; read after the mutation (end_cv)
            core._cv_sharded = True
            cv = None
        else:
            cv = k.volume_variation(ens.u, w, ens.n_total, ens.n_dim)     # reweight.py:417-419
        logz = float(stats[4].item()) if self._logz_host is None else self._logz_host
        st.update_current({"logz": logz, "beta": float(beta), "ess": float(ess), "cv": cv})
        return w

    def _device_search(self, beta_prev: float, target: float):
        """ESS mode: the whole bracket + bisection in one cooperative launch (tb_next_beta)."""
        core = self.core
        ens = core.ensemble
        k = core.k
        flags = 0
        if core.warmup_regime:
            # all stored generations are at beta = 0: every weight is exactly equal and the
            # `ESS <= target` branch depends on numpy's rounding
            ess0 = uniform_weights_ess(ens.n_total_global)
            self.probe_log.append((beta_prev, ess0))
            if ess0 <= target:
                out = k.probe(ens, beta_prev)
                return beta_prev, ess0, out
            flags = 1
        if ens.n_total <= self.SPECULATIVE_MAX:
            # short (local) history: a pass is latency-bound (block merges + synchronisation point >> streaming time), so
            # evaluating the current beta and both possible successors per pass halves the passes for free; on long
            # histories the three-beta pass is fp64-bound and slower (profiles/r02_next_beta_speculative.txt)
            flags |= 2
        timing = core.kernel_timing
        if timing is not None:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        res, plog = k.next_beta(ens, beta_prev, target, flags)
        if timing is not None:
            ev1.record()
        h = res.cpu().numpy()
        nprobe, npass = int(h[6]), int(h[9])     # probes consumed by the search / passes over the ensemble (3 betas each)
        if timing is not None:           # (events, particle-passes of this launch, particle-probes)
            timing.setdefault("next_beta", []).append((ev0, ev1, float(npass) * ens.n_total))
            timing.setdefault("next_beta_probes", []).append((ev0, ev1, float(nprobe) * ens.n_total))
        hp = plog[: 2 * min(nprobe, 512)].cpu().numpy().reshape(-1, 2)
        self.probe_log.extend((float(b), float(e)) for b, e in hp)
        if k.sharded:
            k.consume_exchanges(npass)               # one peer-memory exchange per pass
        if h[8] != 0.0:
            raise FloatingPointError(f"{int(h[8])} non-finite log-weights in the persistent ensemble")
        self._logz_host = float(h[5])                  # logZ(beta) of the last probe is already on the host
        return float(h[0]), float(h[4]), res[1:7]     # (m, S1, S2, ESS, logZ, .) like tb_probe's out


class Trainer:
    """steps/train.py:65-127 + modes.py:131-288 + student.py:6-116."""

    def __init__(self, core):
        self.core = weakref.proxy(core)      # the core owns the steps; no reference cycle, so memory frees on del

    def run(self, weights: Optional[torch.Tensor]) -> ModeStats:
        core = self.core
        ens = core.ensemble
        d = ens.n_dim
        if float(core.state.raw("beta")) == 0.0:      # train.py:79-88
            return ModeStats.identity(d, core.device)
        k = core.k
        lib = k.lib
        st = stream_ptr()
        core._stage("train:trim")
        hook = (lambda: core.resampler.begin_async(weights)) if core.overlap else None
        idx, wt = k.trim(weights, ens.n_total, after_normalize=hook, n_global=ens.n_total_global)
        core._stage("train:draws")
        core.trace["trim_idx"], core.trace["trim_w"] = idx, wt
        n_trim = int(idx.numel())                       # local; equals the global count on one GPU
        n_trim_glob = int(k.last_trim["n_trim"])
        m_total = 4 * n_trim_glob
        if core.config.clustering:
            return self._clustered(idx, wt, n_trim)
        # modes.py:266: renormalise the trimmed weights
        k.g_normalize_begin(wt, n_trim)                 # the sums are not needed on the host: no round trip
        draws = core.rng.train_u(m_total)
        mean, cov, chol, inv = self._fit_mode(idx, wt, n_trim, n_trim_glob, draws, "train_draw_idx")
        dof = torch.full((1,), DOF_FALLBACK, dtype=F64, device=core.device)
        return ModeStats(mean, cov, chol, inv, dof)

    def _clustered(self, idx: torch.Tensor, wt: torch.Tensor, n_trim: int) -> ModeStats:
        """train.py:97-115: (re)fit the hierarchy on the trimmed set, label it, and fit one Student-t
        mode per distinct predicted label (modes.py:131-219)."""
        core = self.core
        ens = core.ensemble
        k = core.k
        cfg = core.config
        d = ens.n_dim
        it = int(core.state.raw("iter"))
        core._stage("train:cluster")
        if it % cfg.cluster_every == 0 or it == 0:                  # train.py:97-104
            core.clusterer.fit(ens.u, wt, rows=idx)
        labels = core.clusterer.predict(ens.u, rows=idx)           # train.py:101 / 112
        core.trace["train_labels"] = labels
        core._stage("train:modes")
        k.g_normalize(wt, n_trim)                                  # modes.py:183
        present = torch.unique(labels).cpu().numpy()               # modes.py:188 (sorted distinct labels)
        draws = core.rng.train_u(4 * n_trim)
        means, covs, chols, invs = [], [], [], []
        offset = 0
        draw_idx = []
        for label in present:
            pos = torch.nonzero(labels == int(label)).reshape(-1)  # modes.py:191, member order kept
            n_c = int(pos.numel())
            rows_c = idx[pos]
            wt_c = torch.empty(n_c, dtype=F64, device=core.device)
            _lib.check(k.lib.tb_take(ptr(wt), ptr(pos), n_c, ptr(wt_c), stream_ptr()), "tb_take")
            k.g_normalize(wt_c, n_c)                               # modes.py:193-194
            mean, cov, chol, inv = self._fit_mode(rows_c, wt_c, n_c, n_c, draws[offset: offset + 4 * n_c], None)
            draw_idx.append(core.trace.get("_last_draw_idx"))
            offset += 4 * n_c
            means.append(mean)
            covs.append(cov)
            chols.append(chol)
            invs.append(inv)
        core.trace["train_draw_idx"] = draw_idx
        K = len(present)
        dof = torch.full((K,), DOF_FALLBACK, dtype=F64, device=core.device)
        return ModeStats(torch.cat(means), torch.cat(covs), torch.cat(chols), torch.cat(invs), dof)

    def _fit_mode(self, idx: torch.Tensor, wt: torch.Tensor, n_trim: int, n_trim_glob: int, draws: torch.Tensor,
                  trace_key: Optional[str]):
        """modes.py:197-209 / 272-282 for one mode: 4n weighted draws of the rows ``idx`` (normalised
        weights ``wt``), then the Student-t fit of student.py:62-94 (median, Sigma, nu = inf)."""
        core = self.core
        ens = core.ensemble
        d = ens.n_dim
        k = core.k
        lib = k.lib
        st = stream_ptr()
        m_total = 4 * n_trim_glob
        didx = k.ws.i64("train_didx", m_total) if trace_key else torch.empty(m_total, dtype=torch.int64,
                                                                              device=core.device)
        if k.sharded:
            # segments of the trimmed set per generation (trim indices are ascending = generation-major)
            seg_begin = torch.searchsorted(idx, core.generation_bounds())
            k.sharded_search(wt, n_trim, seg_begin, draws, didx, "train_cdf", n_global=n_trim_glob)
        else:
            cdf = k.cdf(wt, n_trim, "train_cdf")
            k.search_right(cdf, n_trim, draws, didx)
        counts = k.ws.i32("train_counts", max(n_trim, 1))
        _lib.check(lib.tb_count_indices(ptr(didx), m_total, ptr(counts), max(n_trim, 1), st), "tb_count_indices")
        if trace_key:
            core.trace[trace_key] = didx
        else:
            core.trace["_last_draw_idx"] = didx
        # student.py:62: per-dimension median of the 4n-row multiset (even count: mean of the middle pair)
        core._stage("train:median")
        pair = k.ws.f64("train_pair", 2 * d)
        done = False
        if not k.sharded:
            mws_ = k.ws.bytes("unit_median", lib.tb_unit_median_workspace_bytes(d))
            ovf = k.ws.i32("median_ovf", 1)
            _lib.check(lib.tb_unit_median_pair(ptr(ens.u), ptr(idx), ptr(counts), n_trim, d, m_total // 2 - 1,
                                               ptr(mws_), ptr(pair), ptr(ovf), st), "tb_unit_median_pair")
            done = int(ovf.item()) == 0
        else:
            done = k.bucket_pair_sharded(ens.u, idx if n_trim else None, counts if n_trim else None, n_trim, d,
                                         m_total // 2 - 1, False, 0.0, 1.0, True, True, pair)
        if not done:
            k.g_select_pair(ens.u, idx, d, n_trim, d, counts, m_total // 2 - 1, False, pair)
        mean = torch.empty((1, d), dtype=F64, device=core.device)
        _lib.check(lib.tb_median_pairs(ptr(pair), d, ptr(mean), st), "tb_median_pairs")
        # student.py:63: Sigma = cov(ddof=1)*(M-1)/M + diag(var)/M from count-weighted moments
        core._stage("train:moments")
        mws = k.ws.bytes("mom", lib.tb_moments_workspace_bytes(d))
        cmean = k.ws.f64("train_cmean", d)
        scatter = k.ws.f64("train_scatter", d * d)
        if k.sharded:
            from .sharded import sharded_mode_moments

            sharded_mode_moments(core, idx, counts, n_trim, m_total, cmean, scatter)
        else:
            _lib.check(lib.tb_counted_moments(ptr(ens.u), ptr(idx), ptr(counts), n_trim, d, 1.0 / m_total, ptr(mws),
                                              ptr(cmean), ptr(scatter), st), "tb_counted_moments")
        cov = torch.empty((1, d, d), dtype=F64, device=core.device)
        _lib.check(lib.tb_student_sigma(ptr(scatter), d, float(m_total), ptr(cov), st), "tb_student_sigma")
        # student.py:75-79 + modes.py:105-119: Cholesky (regularise on failure), inverse.  The EM loop
        # returns on its first pass with nu = inf (student.py:54-55,93-94) -> dof = 1e6.
        chol = torch.empty((1, d, d), dtype=F64, device=core.device)
        inv = torch.empty((1, d, d), dtype=F64, device=core.device)
        info = k.ws.i32("train_info", 1)
        _lib.check(lib.tb_chol_inv(ptr(cov), d, 1, ptr(chol), ptr(inv), ptr(info), None, st), "tb_chol_inv")
        return mean, cov, chol, inv


class Resampler:
    """steps/resample.py:52-99."""

    def __init__(self, core):
        self.core = weakref.proxy(core)      # the core owns the steps; no reference cycle, so memory frees on del
        self._pending = None

    def begin_async(self, weights: torch.Tensor) -> None:
        """Multinomial resampling on the side stream, issued as soon as trim_weights has normalised the
        weights in place (tools.py:36) -- from then on they are read-only -- so the cumulative sum, the
        search and the row gather overlap with the rest of the Trainer on the main stream."""
        core = self.core
        ens = core.ensemble
        n, d = core.n_local, ens.n_dim
        ready = torch.cuda.Event()
        ready.record()
        u = torch.empty((n, d), dtype=F64, device=core.device)          # allocated (and later consumed) on the main stream
        logl = torch.empty(n, dtype=F64, device=core.device)
        core.side.wait_event(ready)
        if core.comm.on:
            from .sharded import sharded_resample

            with torch.cuda.stream(core.side):
                ks = core.k_side
                ks.ws.hint = ens.cap
                draws = core.rng.resample_u(core.n_global)
                u, logl = sharded_resample(core, weights, draws, k=ks)     # views into ks's peer-mapped row buffers
                done = torch.cuda.Event()
                done.record()
            self._pending = (u, logl, core.trace.get("resample_idx"), done)
            return
        with torch.cuda.stream(core.side):
            ks = core.k_side
            cdf = ks.cdf(weights, ens.n_total)
            idx = ks.ws.i64("res_idx", n)
            ks.search_right(cdf, ens.n_total, core.rng.resample_u(n), idx)
            _lib.check(ks.lib.tb_gather_rows(ptr(ens.u), ptr(ens.logl), d, ptr(idx), n, ptr(u), ptr(logl),
                                             stream_ptr()), "tb_gather_rows")
            done = torch.cuda.Event()
            done.record()
        self._pending = (u, logl, idx, done)

    def run(self, weights: Optional[torch.Tensor]) -> None:
        core = self.core
        st = core.state
        n = core.n_local
        if float(st.raw("beta")) == 0.0:               # resample.py:69-72
            st.set_current("assignments", core.zero_assignments(n))
            return
        ens = core.ensemble
        k = core.k
        if self._pending is not None:
            u, logl, idx, done = self._pending
            self._pending = None
            torch.cuda.current_stream().wait_event(done)
            core.trace["resample_idx"] = idx
            core.assign = None
            st.update_current({"u": u, "x": None, "logl": logl, "assignments": core.zero_assignments(n)})
            return
        if k.sharded:
            from .sharded import sharded_resample

            if core.config.resample == "mult":
                u, logl = sharded_resample(core, weights, core.rng.resample_u(core.n_global))
            else:
                u, logl = sharded_resample(core, weights, None, systematic=True, u0=core.rng.resample_u0())
            st.update_current({"u": u, "x": None, "logl": logl, "assignments": core.zero_assignments(n)})
            return
        cdf = k.cdf(weights, ens.n_total)
        idx = k.ws.i64("res_idx", n)
        if core.config.resample == "mult":
            k.search_right(cdf, ens.n_total, core.rng.resample_u(n), idx)
        else:
            k.systematic(cdf, ens.n_total, core.rng.resample_u0(), n, idx)
        core.trace["resample_idx"] = idx
        u = torch.empty((n, ens.n_dim), dtype=F64, device=core.device)
        logl = torch.empty(n, dtype=F64, device=core.device)
        _lib.check(k.lib.tb_gather_rows(ptr(ens.u), ptr(ens.logl), ens.n_dim, ptr(idx), n, ptr(u), ptr(logl),
                                        stream_ptr()), "tb_gather_rows")
        if core.config.clustering:                     # resample.py:92-94
            assign = core.clusterer.predict(u)
            core.assign = assign
            core.trace["assignments"] = assign
            st.update_current({"u": u, "x": None, "logl": logl, "assignments": assign})
        else:
            core.assign = None
            st.update_current({"u": u, "x": None, "logl": logl, "assignments": core.zero_assignments(n)})


class Mutator:
    """steps/mutate.py:76-200 + mcmc.py:142-323."""

    CHUNK = 8   # Metropolis steps enqueued between looks at the device-side stop flag

    def _external_loop(self, params, tape, tape_ref, assign, u, logl, qcur, ws, ctrl, n_cap: int):
        """Split Metropolis steps around the user's callables (mcmc.py:142-208): proposal kernel ->
        x = prior_transform(u'), logl' = log_likelihood(x') -> accept kernel with sigma adaptation and
        the stop rule.  Device callables are enqueued without a host round trip and the stop flag is read
        every CHUNK steps; numpy callables synchronise every step anyway."""
        core = self.core
        k = core.k
        lib = k.lib
        sp = stream_ptr()
        n, d = core.n_local, core.config.n_dim
        u_prop = k.ws.f64("mcmc_u_prop", n * d).reshape(n, d)
        meta = k.ws.i32("mcmc_meta", n)
        on_device = (core.bridge.prior_registry or core.bridge.prior_device) and core.bridge.like_device
        look_every = self.CHUNK if on_device else 1
        launched = 0
        while True:
            if tape is not None and launched >= tape.steps:
                raise RuntimeError(
                    f"tape holds {tape.steps} MCMC steps but the device stop rule has not fired after {launched}")
            _lib.check(lib.tb_mcmc_propose(n, C.byref(params), tape_ref, ptr(assign), ptr(u), ptr(logl), ptr(qcur),
                                           ptr(ws), ptr(ctrl), ptr(u_prop), ptr(meta), sp), "tb_mcmc_propose")
            logl_prop = core.bridge.like(core.bridge.prior(u_prop, core))
            _lib.check(lib.tb_mcmc_accept(n, C.byref(params), tape_ref, ptr(assign), ptr(u), ptr(logl), ptr(qcur),
                                          ptr(ws), ptr(ctrl), ptr(u_prop), ptr(logl_prop), ptr(meta), sp),
                       "tb_mcmc_accept")
            launched += 1
            if launched % look_every == 0 or launched >= n_cap:
                h = ctrl.cpu().numpy()
                if h[1] != 0.0 or launched >= n_cap:
                    return h, launched

    def __init__(self, core):
        self.core = weakref.proxy(core)      # the core owns the steps; no reference cycle, so memory frees on del

    def run(self, mode_stats: ModeStats) -> None:
        core = self.core
        st = core.state
        cfg = core.config
        n, d = core.n_local, cfg.n_dim
        k = core.k
        lib = k.lib
        sp = stream_ptr()
        beta = float(st.raw("beta"))
        params = core.mcmc_params(beta, mode_stats)
        if beta == 0.0:                                # mutate.py:100-149
            u = torch.empty((n, d), dtype=F64, device=core.device)
            logl = torch.empty(n, dtype=F64, device=core.device)
            tape_u = core.rng.prior_u(n, d)
            if core.bridge.external:                   # same uniforms; prior / likelihood by the user's callables
                _lib.check(lib.tb_prior_draw(n, C.byref(params), ptr(tape_u), ptr(u), None, None, sp), "tb_prior_draw")
                logl = core.bridge.like(core.bridge.prior(u, core))
            else:
                _lib.check(lib.tb_prior_draw(n, C.byref(params), ptr(tape_u), ptr(u), None, ptr(logl), sp),
                           "tb_prior_draw")
            st.update_current({"u": u, "x": None, "logl": logl, "assignments": core.zero_assignments(n),
                               "calls": st.raw("calls") + core.n_global, "steps": 1, "acceptance": 1.0,
                               "efficiency": 1.0})
            bad = torch.isinf(logl)
            any_bad = bool(bad.any())
            if k.sharded:
                any_bad = bool(k.g_int(int(any_bad)))
                if any_bad:
                    # rare: replicate the generation (rank-major = slot order), apply mutate.py:122-148 to the
                    # replicated table with the replicated picks, keep this rank's block of slots
                    U = k.comm.allgather(u).reshape(core.n_global, d)
                    L = k.comm.allgather(logl).reshape(core.n_global)
                    badg = torch.isinf(L)
                    every = torch.arange(core.n_global, device=core.device)
                    inf_idx, fin_idx = every[badg], every[~badg]
                    if fin_idx.numel() > 0:
                        pick = core.rng.inf_pick(fin_idx, int(inf_idx.numel()))
                        U[inf_idx] = U[pick]
                        L[inf_idx] = L[pick]
                    lo = core.slot_offset
                    u.copy_(U[lo:lo + n])
                    logl.copy_(L[lo:lo + n])
                    st.set_current("logz", st.raw("logz") + math.log(fin_idx.numel() / core.n_global))
                    return
            if any_bad:                                 # mutate.py:122-148 (rare; bookkeeping on device tensors)
                every = torch.arange(n, device=core.device)
                inf_idx, fin_idx = every[bad], every[~bad]
                if fin_idx.numel() > 0:
                    pick = core.rng.inf_pick(fin_idx, int(inf_idx.numel()))
                    u[inf_idx] = u[pick]
                    logl[inf_idx] = logl[pick]
                st.set_current("logz", st.raw("logz") + math.log(fin_idx.numel() / n))
            return
        u = st.raw("u")
        logl = st.raw("logl")
        K = mode_stats.K
        assign = core.assign if cfg.clustering else None
        if assign is not None and int(assign.max().item()) >= K:
            # modes.py:188: modes are the distinct labels of the trimmed set; a walker labelled beyond
            # them indexes past the mode arrays in the reference too (mcmc.py:228-231)
            raise IndexError(f"walker assigned to cluster {int(assign.max().item())} but only {K} modes were fitted")
        ctrl = k.ws.f64("mcmc_ctrl", int(lib.tb_mcmc_ctrl_doubles(K)))
        ws = k.ws.bytes("mcmc_ws", lib.tb_mcmc_workspace_bytes(n, K))
        qcur = k.ws.f64("mcmc_q", 2 * n)                 # q_k and the cached Student-t term of the current state
        _lib.check(lib.tb_mcmc_begin(n, C.byref(params), ptr(assign), ptr(u), ptr(qcur), ptr(ws), ptr(ctrl), sp),
                   "tb_mcmc_begin")
        tape = core.rng.mcmc_tape(n, d)
        tape_ref = C.byref(tape) if tape is not None else None
        n_min = cfg.n_steps * d
        n_cap = cfg.n_max_steps * d
        fused = k.sharded and k.xgpu is not None and K + 3 <= 71
        if core.bridge.external:
            h, launched = self._external_loop(params, tape, tape_ref, assign, u, logl, qcur, ws, ctrl, n_cap)
        elif k.sharded:
            k.comm.allreduce_sum_(ctrl[8 + K: 8 + 2 * K])          # walkers per mode: global counts
        if k.sharded and not fused and not core.bridge.external:
            from .sharded import sharded_mcmc_loop

            h, launched = sharded_mcmc_loop(core, params, tape_ref, u, logl, qcur, ws, ctrl, min(n_min, n_cap),
                                            n_cap, self.CHUNK)
        elif not core.bridge.external:
            # ONE persistent launch runs the whole mutation: the stop rule (mcmc.py:104-135, 192-194) is evaluated
            # on the device every step, and on > 1 GPU the per-step totals travel over peer memory inside the kernel
            if fused:
                params.xgpu = C.pointer(k.xgpu)
                params.defer_update = 0
            budget = n_cap if tape is None else min(n_cap, tape.steps)
            timing = core.kernel_timing
            if timing is not None:
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
            _lib.check(lib.tb_mcmc_steps(n, C.byref(params), tape_ref, ptr(assign), ptr(u), ptr(logl), ptr(qcur),
                                         ptr(ws), ptr(ctrl), int(budget), sp), "tb_mcmc_steps")
            if timing is not None:
                ev1.record()
            h = ctrl.cpu().numpy()
            if timing is not None:       # (events, walker-steps of this launch): bench.py's roofline of the dominant kernel
                timing.setdefault("mcmc", []).append((ev0, ev1, float(h[0]) * n))
            launched = 1
            if fused:
                k.consume_exchanges(int(h[0]))                     # one exchange per executed step
            if h[1] == 0.0 and int(h[0]) < n_cap:
                raise RuntimeError(
                    f"tape holds {tape.steps if tape is not None else 0} MCMC steps but the device stop rule has "
                    f"not fired after {int(h[0])}")
        core.n_mcmc_launches += launched
        if h[4] != 0.0:
            raise RuntimeError(f"MCMC kernel error code {int(h[4])} (1: tape exhausted, 2: proposal never entered the "
                               "cube, 3: a peer GPU did not answer)")
        steps = int(h[0])
        sig = h[8:8 + K]
        sigma0 = 2.38 / math.sqrt(d)
        core.trace["mcmc_sigma"] = sig.copy()
        st.update_current({"u": u, "x": None, "logl": logl, "efficiency": float(np.mean(sig) / sigma0),
                           "acceptance": float(h[3]), "steps": steps,
                           "calls": st.raw("calls") + steps * core.n_global})
