"""Multi-GPU (sharded ensemble) variants of the device steps: the same kernels on each rank's
shard, with the handful of global reductions of SURVEY 8e carried by NCCL (see dist.py).

Results do not depend on the number of GPUs beyond fp64 summation order: Philox counters are keyed
by GLOBAL walker slot, the uniforms of the resampling / training draws are replicated, and every
decision is taken from all-reduced quantities that are bitwise identical on all ranks.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import numpy as np
import torch

from . import _lib
from .dist import Comm, exchange_owned_rows, merge_ess_triples
from .ensemble import PersistentEnsemble, ptr, stream_ptr
from .steps import _EPS, F64, Kernels


# Symmetric (peer-mapped) allocations are expensive to set up -- cuMemCreate + handle exchange + mapping on every rank,
# 30-90 ms for the four a sharded Sampler needs -- and carry only transient state (sequence-numbered flags), so they are
# pooled per process: a Sampler leases them and hands them back when it is garbage-collected; the next Sampler of the
# same shape reuses them with the sequence numbers simply continuing.  Every rank creates and drops Samplers in the same
# order, so the pools stay aligned across ranks.
_SYMM_POOL: dict = {}


def _pool_acquire(key, factory, fits=None):
    idle = _SYMM_POOL.setdefault(key, [])
    for i, obj in enumerate(idle):
        if fits is None or fits(obj):
            return idle.pop(i)
    return factory()


def _pool_release(leased) -> None:
    for key, obj in leased:
        _SYMM_POOL.setdefault(key, []).append(obj)
    leased.clear()


class PeerCollectives:
    """All-reduce (sum) / all-gather of small tensors through peer-mapped staging buffers (csrc/tb_xcoll.cu): two
    kernels on the current stream per call, no host synchronisation, bitwise identical results on every rank
    (payloads are folded in rank order).  Every rank must issue the same sequence of calls."""

    CAP_BYTES = 4 << 20
    DTYPES = {torch.float64: 0, torch.int64: 1, torch.int32: 2}

    def __init__(self, lib, device: torch.device, comm: Comm):
        import torch.distributed._symmetric_memory as symm

        self.lib, self.device, self.comm = lib, device, comm
        n = int(lib.tb_xcoll_buffer_bytes(self.CAP_BYTES)) // 8
        self.buf = symm.empty(n, dtype=F64, device=device)
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, torch.distributed.group.WORLD if comm.group is None else comm.group)
        torch.cuda.synchronize()
        torch.distributed.all_reduce(torch.zeros(1, device=device))          # everyone has zeroed its staging buffer
        self.ticket = torch.zeros(4, dtype=torch.int32, device=device)
        self.err = torch.zeros(1, dtype=torch.int32, device=device)
        self.x = _lib.TbXcoll()
        self.x.rank, self.x.world, self.x.seq, self.x.cap_bytes = comm.rank, comm.world, 0, self.CAP_BYTES
        self.x.ticket = self.ticket.data_ptr()
        for r in range(comm.world):
            self.x.peer[r] = int(self.handle.buffer_ptrs[r])

    def takes(self, t: torch.Tensor) -> bool:
        return (t.is_cuda and t.is_contiguous() and t.numel() > 0 and t.element_size() >= 4
                and t.numel() * t.element_size() <= self.CAP_BYTES)

    def allreduce_sum_(self, t: torch.Tensor) -> torch.Tensor:
        code = self.DTYPES.get(t.dtype)
        if code is None:
            torch.distributed.all_reduce(t, group=self.comm.group)
            return t
        self.x.seq += 1
        _lib.check(self.lib.tb_xcoll_allreduce_sum(ptr(t), t.numel(), code, C.byref(self.x), ptr(self.err), stream_ptr()),
                   "tb_xcoll_allreduce_sum")
        return t

    def allgather(self, flat: torch.Tensor) -> torch.Tensor:
        out = torch.empty(self.comm.world * flat.numel(), dtype=flat.dtype, device=flat.device)
        self.x.seq += 1
        _lib.check(self.lib.tb_xcoll_allgather(ptr(flat), flat.numel() * flat.element_size(), ptr(out), C.byref(self.x),
                                               ptr(self.err), stream_ptr()), "tb_xcoll_allgather")
        return out

    def check(self) -> None:
        if int(self.err.item()):
            raise RuntimeError("peer-memory collective timed out: a peer GPU did not answer")


class ShardedKernels(Kernels):
    BUCKET_SLOTS = 1024      # candidate slots per column and rank gathered by the sharded bucket select
    def __init__(self, device: torch.device, comm: Comm, role: str = "main"):
        super().__init__(device, role)
        import weakref

        self.comm = comm
        self.sharded = True
        self._leased = []                                   # (pool key, object) pairs, returned when this object dies
        weakref.finalize(self, _pool_release, self._leased)
        self._pool_tag = (device.index, comm.world, id(comm.group) if comm.group is not None else 0, role)
        self.xgpu = self._setup_xgpu()
        if self.xgpu is not None and comm.fast is None:
            import os

            if os.environ.get("TEMPEST_B200_PEER_COLLECTIVES", "1") != "0":
                key = ("coll",) + self._pool_tag
                comm.fast = _pool_acquire(key, lambda: PeerCollectives(self.lib, device, comm))
                self._leased.append((key, comm.fast))

    # -- peer-mapped exchange buffers for the in-kernel collectives (tb_xgpu.cuh) --------------------
    def _setup_xgpu(self):
        """Symmetric memory rendezvous: every rank learns the device address of every peer's exchange
        buffer.  Returns None (NCCL collectives between launches are used instead) when the platform
        cannot map peer memory."""
        if self.comm.world > 8:
            return None

        def make():
            import torch.distributed._symmetric_memory as symm

            n = int(self.lib.tb_xgpu_buffer_bytes()) // 8
            buf = symm.empty(n, dtype=F64, device=self.device)
            buf.zero_()
            handle = symm.rendezvous(buf, torch.distributed.group.WORLD if self.comm.group is None else self.comm.group)
            torch.cuda.synchronize()
            self.comm.allreduce_sum_(torch.zeros(1, device=self.device))      # everyone has zeroed its buffer
            x = _lib.TbXgpu()
            x.rank, x.world, x.seq = self.comm.rank, self.comm.world, 1
            for r in range(self.comm.world):
                x.peer[r] = int(handle.buffer_ptrs[r])
            return (buf, handle, x)

        try:
            key = ("xgpu",) + self._pool_tag
            obj = _pool_acquire(key, make)
            self._leased.append((key, obj))
            self._xbuf, self._xhandle, x = obj
            return x                  # the sequence number continues where the previous lessee stopped
        except Exception as exc:      # pragma: no cover - platform dependent
            import warnings

            warnings.warn(f"peer-mapped exchange buffers unavailable ({type(exc).__name__}: {exc}); "
                          "falling back to NCCL collectives between launches")
            return None

    def consume_exchanges(self, count: int) -> None:
        self.xgpu.seq += int(count)

    def next_beta(self, ens, beta_prev: float, target: float, flags: int, log_cap: int = 512):
        """Whole ESS bracket + bisection in one cooperative launch per rank; the per-probe merge of the
        ranks' (m, S1, S2) runs inside the kernel over NVLink peer memory."""
        res = self.ws.f64("nb_res", 16)
        plog = self.ws.f64("nb_log", 2 * log_cap)
        _lib.check(self.lib.tb_next_beta_x(ptr(ens.logl), ptr(ens.C), ens.n_total, float(beta_prev), float(target),
                                           int(flags), ptr(self._probe_ws), ptr(res), ptr(plog), log_cap,
                                           C.byref(self.xgpu), stream_ptr()), "tb_next_beta_x")
        return res, plog

    # -- reweighting: merge the per-shard (m, S1, S2) triples in rank order -------------------------
    def probe(self, ens: PersistentEnsemble, beta: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        out = super().probe(ens, beta, out)
        allp = self.comm.allgather(out[:6]).cpu().numpy()
        m, s1, s2 = merge_ess_triples(allp[:, :3])
        merged = [m, s1, s2, s1 * s1 / s2, m + math.log(s1), float(allp[:, 5].sum())]
        out[:6].copy_(torch.tensor(merged, dtype=F64))
        return out

    # -- hooks used by Kernels.trim -------------------------------------------------------------------
    def g_int(self, value: int) -> int:
        t = torch.tensor([int(value)], dtype=torch.int64, device=self.device)
        return int(self.comm.allreduce_sum_(t).item())

    def g_sum3(self, vals, m: int, thr: float):
        out3 = self.ws.f64("trim_m3", 3)
        if m > 0:
            _lib.check(self.lib.tb_masked_sums(ptr(vals), m, float(thr), ptr(self._reduce_ws), ptr(out3),
                                               stream_ptr()), "tb_masked_sums")
        else:
            out3.zero_()
        c, a1, a2 = self.comm.allreduce_sum_(out3).cpu().numpy()
        return float(c), float(a1), float(a2)

    def g_normalize_begin(self, w, n: int):
        """w /= global sum(w), enqueued only: local sums -> peer-memory all-reduce -> scale by the device scalar."""
        out3 = self.ws.f64("norm_m3", 3)
        if n > 0:
            _lib.check(self.lib.tb_masked_sums(ptr(w), n, -math.inf, ptr(self._reduce_ws), ptr(out3), stream_ptr()),
                       "tb_masked_sums")
        else:
            out3.zero_()                         # a rank without elements still takes part in the collective
        self.comm.allreduce_sum_(out3)
        if n > 0:
            _lib.check(self.lib.tb_scale_inplace_dev(ptr(w), n, ptr(out3[1:]), stream_ptr()), "tb_scale_inplace_dev")
        return out3

    def g_normalize_end(self, out3):
        _, s1, s2 = out3.cpu().numpy()
        return float(s1), float(s2) / (float(s1) * float(s1))

    def g_hist(self, w, n: int):
        buf, cnt, s1, s2 = self._hist_buffers("trim_hist")
        if n > 0:
            _lib.check(self.lib.tb_binade_hist(ptr(w), n, ptr(cnt), ptr(s1), ptr(s2), stream_ptr()), "tb_binade_hist")
        else:
            buf.zero_()
        self.comm.allreduce_sum_(cnt)            # counts (int64) and the two sum columns (fp64, contiguous): two
        self.comm.allreduce_sum_(buf[2048:])     # collectives, one read-back
        return self._hist_to_host(buf)

    def g_subhist(self, w, n: int, binade: int):
        buf, cnt, s1, s2 = self._hist_buffers("trim_hist2")
        if n > 0:
            _lib.check(self.lib.tb_subbin_hist(ptr(w), n, int(binade), ptr(cnt), ptr(s1), ptr(s2), stream_ptr()),
                       "tb_subbin_hist")
        else:
            buf.zero_()
        self.comm.allreduce_sum_(cnt)
        self.comm.allreduce_sum_(buf[2048:])
        return self._hist_to_host(buf)

    def bucket_pair_sharded(self, base, rows, mult, n: int, d: int, rank_lo: int, same: bool, lo_value: float,
                            hi_value: float, unit_map: bool, build: bool, out: torch.Tensor) -> bool:
        """Order statistics (rank_lo, rank_lo + 1) of every column of the GLOBAL multiset with the bucket
        method: local 65 536-bucket histograms are all-reduced, every rank picks the same bucket(s), compacts
        its own candidates, and the (few hundred) candidates are merged on every rank before the exact
        in-block select.  Two small collectives instead of one all-reduce per radix level.  Returns False when
        a bucket holds more candidates than the in-block select takes (the caller falls back to g_select_pair)."""
        lib = self.lib
        st = stream_ptr()
        G = self.comm.world
        bws = self.ws.bytes(f"bucket_sh{d}", lib.tb_unit_median_workspace_bytes(d))
        if not hasattr(self, "_bucket_off"):
            self._bucket_off = {}
        if d not in self._bucket_off:
            off = np.zeros(4, dtype=np.int64)
            _lib.check(lib.tb_bucket_offsets(d, off.ctypes.data), "tb_bucket_offsets")
            self._bucket_off[d] = [int(v) for v in off]
        o0, o1, o2, o3 = self._bucket_off[d]
        cap = 65536
        ovf = self.ws.i32("bucket_ovf", 1)

        def stage(k, out_=None, ovf_=None):
            _lib.check(lib.tb_bucket_stage(ptr(base), ptr(rows), ptr(mult), n, d, int(rank_lo), int(same), float(lo_value),
                                           float(hi_value), int(unit_map), k, ptr(bws), ptr(out_), ptr(ovf_), st),
                       "tb_bucket_stage")

        if build:
            stage(0)
            self.comm.allreduce_sum_(bws[o0: o0 + 4 * d * cap].view(torch.int32))
        stage(1)
        stage(2)
        # merge the ranks' candidate lists on the device: a fixed number of slots per column and rank is gathered, so no
        # count has to visit the host before the exact select (one read -- the overflow flag -- per call instead of two
        # reads and a dozen eager tensor ops; a column with more candidates than slots takes the fallback of the caller)
        capx = self.BUCKET_SLOTS
        sel = bws[o1: o1 + 24 * d].view(torch.int32).reshape(d, 6)
        cval = bws[o2: o2 + 8 * d * cap].view(F64).reshape(d, cap)
        cmul = bws[o3: o3 + 4 * d * cap].view(torch.int32).reshape(d, cap)
        gsel = self.comm.allgather(sel[:, 4:6].contiguous())                      # [G, d, (count, overflow)]
        gv = self.comm.allgather(cval[:, :capx].contiguous())                     # [G, d, capx]
        gm = self.comm.allgather(cmul[:, :capx].contiguous())
        _lib.check(lib.tb_bucket_merge(ptr(gsel), ptr(gv), ptr(gm), G, d, capx, ptr(bws), st), "tb_bucket_merge")
        stage(3, out, ovf)
        return int(ovf.item()) == 0

    def g_select(self, base, rows, stride, m, ncols, mult, ranks, nranks, out):
        """Distributed radix select: local histograms, all-reduced per level, replicated picks."""
        lib = self.lib
        sws = self.ws.bytes("select", lib.tb_select_workspace_bytes(ncols, nranks))
        off = int(lib.tb_select_hist_offset(ncols, nranks))
        hist = sws[off: off + 4 * ncols * nranks * 2048].view(torch.int32)
        args = (ptr(base), ptr(rows), stride, m, ncols, ptr(mult), ptr(ranks), nranks, ptr(sws), ptr(out))
        st = stream_ptr()
        _lib.check(lib.tb_select_stage(*args, 0, 0, st), "tb_select_stage")
        for level in range(6):
            _lib.check(lib.tb_select_stage(*args, 1, level, st), "tb_select_stage")
            self.comm.allreduce_sum_(hist)
            _lib.check(lib.tb_select_stage(*args, 2, level, st), "tb_select_stage")
        _lib.check(lib.tb_select_stage(*args, 3, 0, st), "tb_select_stage")
        return out

    def g_select_pair(self, base, rows, stride, m, ncols, mult, rank_lo, same, out):
        ranks = torch.tensor([rank_lo, rank_lo if same else rank_lo + 1], dtype=torch.int64, device=self.device)
        return self.g_select(base, rows, stride, m, ncols, mult, ranks, 2, out)

    # -- volume variation: partial moments + all-reduce ------------------------------------------------
    def volume_variation(self, u, w, n: int, d: int) -> float:
        lib, st = self.lib, stream_ptr()
        if n * self.comm.world < d + 1:
            return 1e10
        ws = self.ws.bytes("mom", lib.tb_moments_workspace_bytes(d))
        mean = self.ws.f64("vv_mean", d)
        cov = self.ws.f64("vv_cov", d * d)
        _lib.check(lib.tb_moments_partial(ptr(u), None, ptr(w), None, n, d, 1.0, 1, 0, ptr(ws), ptr(mean), None, st),
                   "tb_moments_partial")
        self.comm.allreduce_sum_(mean)
        _lib.check(lib.tb_moments_partial(ptr(u), None, ptr(w), None, n, d, 1.0, 0, 1, ptr(ws), ptr(mean), ptr(cov), st),
                   "tb_moments_partial")
        self.comm.allreduce_sum_(cov)
        work = self.ws.f64("vv_work", d * d)
        inv = self.ws.f64("vv_inv", d * d)
        info = self.ws.i32("vv_info", 1)
        norms = self.ws.f64("vv_norms", 3)
        for attempt in range(2):
            work.copy_(cov)
            _lib.check(lib.tb_chol_inv(ptr(work), d, 1, None, ptr(inv), ptr(info), ptr(norms), st), "tb_chol_inv")
            code = int(info.item())
            nr = norms.cpu().numpy()
            if attempt == 0:
                singular = code != 0 or not np.isfinite(nr[:2]).all() or nr[0] * nr[1] >= 1.0 / (d * _EPS)
            else:
                singular = code == 2 or not np.isfinite(nr[:2]).all()
            if not singular:
                break
            if attempt == 1:
                return 1e10
            _lib.check(lib.tb_add_trace_reg(ptr(cov), d, 1e-6, st), "tb_add_trace_reg")
        out = self.ws.f64("vv_out", 2)
        _lib.check(lib.tb_mahalanobis_cv(ptr(u), ptr(w), n, d, ptr(mean), ptr(inv), ptr(self._reduce_ws), ptr(out), st),
                   "tb_mahalanobis_cv")
        raw = self.comm.allreduce_sum_(out[1:2].clone())
        return 0.5 * math.sqrt(float(raw.item()))

    def volume_variation_async(self, u, w, n: int, d: int) -> None:
        """tools.py:58-117 enqueued without a host round trip (ESS mode: cv is a diagnostic, read after the mutation):
        partial moments + all-reduces, Cholesky / rank decision / regularisation / second inverse on the device."""
        lib, st = self.lib, stream_ptr()
        if n * self.comm.world < d + 1:
            self._vv_async = "small"
            return
        ws = self.ws.bytes("mom", lib.tb_moments_workspace_bytes(d))
        mean, cov = self.ws.f64("vv_mean", d), self.ws.f64("vv_cov", d * d)
        work, inv = self.ws.f64("vv_work", d * d), self.ws.f64("vv_inv", d * d)
        info, info2, flags = self.ws.i32("vv_info", 1), self.ws.i32("vv_info2", 1), self.ws.i32("vv_flags", 4)
        norms, norms2 = self.ws.f64("vv_norms", 3), self.ws.f64("vv_norms2", 3)
        _lib.check(lib.tb_moments_partial(ptr(u), None, ptr(w), None, n, d, 1.0, 1, 0, ptr(ws), ptr(mean), None, st),
                   "tb_moments_partial")
        self.comm.allreduce_sum_(mean)
        _lib.check(lib.tb_moments_partial(ptr(u), None, ptr(w), None, n, d, 1.0, 0, 1, ptr(ws), ptr(mean), ptr(cov), st),
                   "tb_moments_partial")
        self.comm.allreduce_sum_(cov)
        work.copy_(cov)
        _lib.check(lib.tb_chol_inv(ptr(work), d, 1, None, ptr(inv), ptr(info), ptr(norms), st), "tb_chol_inv")
        _lib.check(lib.tb_vv_regularise(ptr(cov), ptr(work), d, ptr(info), ptr(norms), ptr(flags), st), "tb_vv_regularise")
        _lib.check(lib.tb_chol_inv(ptr(work), d, 1, None, ptr(inv), ptr(info2), ptr(norms2), st), "tb_chol_inv")
        out = self.ws.f64("vv_out", 2)
        _lib.check(lib.tb_mahalanobis_cv(ptr(u), ptr(w), n, d, ptr(mean), ptr(inv), ptr(self._reduce_ws), ptr(out), st),
                   "tb_mahalanobis_cv")
        raw = self.ws.f64("vv_raw", 2)
        raw[:1].copy_(out[1:2])
        self.comm.allreduce_sum_(raw[:1])
        res = self.ws.f64("vv_res", 2)
        _lib.check(lib.tb_vv_finish(ptr(raw), ptr(info2), ptr(norms2), ptr(flags), ptr(res), st), "tb_vv_finish")
        self._vv_async = "enqueued"

    def volume_variation_result(self) -> float:
        if self._vv_async == "small":
            return 1e10
        return float(self.ws.f64("vv_res", 2)[0].item())

    # -- global exact cumulative sum + searches over the sharded weight vector (csrc/tb_cdf.cu) ------------------
    def _cdf_tables(self, need_cap: int):
        """Peer-mapped table memory of the sharded cdf (tile sums / classes / totals and the elements of the hard
        tiles travel through it by NVLink stores).  One allocation sized for the ensemble's capacity."""
        if getattr(self, "_cdf_cap", 0) >= need_cap:
            return

        def make():
            import torch.distributed._symmetric_memory as symm

            cap = int(need_cap)
            n = int(self.lib.tb_cdf_x_table_bytes(cap)) // 8
            buf = symm.empty(n, dtype=F64, device=self.device)
            buf.zero_()
            handle = symm.rendezvous(buf, torch.distributed.group.WORLD if self.comm.group is None else self.comm.group)
            torch.cuda.synchronize()
            self.comm.allreduce_sum_(torch.zeros(1, device=self.device))      # everyone has zeroed its tables
            return {"buf": buf, "handle": handle, "cap": cap, "seq": 0}

        key = ("cdf",) + self._pool_tag
        t = _pool_acquire(key, make, fits=lambda o: o["cap"] >= need_cap)
        self._leased.append((key, t))
        self._cdf_tab = t
        self._cdf_buf, self._cdf_handle, self._cdf_cap = t["buf"], t["handle"], t["cap"]

    def cdf_x(self, p: torch.Tensor, n: int, seg_begin: torch.Tensor, n_global: int, name: str):
        """numpy's sequential cumsum of the GLOBAL weight vector (generation-major, rank-minor: the order of the
        single-GPU ensemble), this rank's part in ``cdf``.  ``seg_begin`` [S+1]: local start positions of this rank's
        per-generation segments; ``n_global``: global element count (the same on every rank)."""
        S = int(seg_begin.numel()) - 1
        hint = max(self.ws.hint * self.comm.world, int(n_global), 1)
        self._cdf_tables(int(self.lib.tb_cdf_tile_cap(hint, 64 * self.comm.world)))
        cap = self._cdf_cap
        if int(self.lib.tb_cdf_tile_cap(int(n_global), S * self.comm.world)) > cap:
            raise RuntimeError("sharded cdf: tile table capacity exceeded")
        ws = self.ws.bytes("cdfx_" + name, self.lib.tb_cdf_x_workspace_bytes(cap))
        cdf = self.ws.f64(name, max(n, 1))
        self._cdf_tab["seq"] += 1        # per table memory, continuing across the Samplers that lease it
        x = _lib.TbCdfX()
        x.rank, x.world, x.seq = self.comm.rank, self.comm.world, self._cdf_tab["seq"]
        for r in range(self.comm.world):
            x.peer[r] = int(self._cdf_handle.buffer_ptrs[r])
        _lib.check(self.lib.tb_cdf_exact_x(ptr(p) if n else None, n, ptr(seg_begin), S, int(n_global), cap, ptr(cdf),
                                           ptr(ws), C.byref(x), stream_ptr()), "tb_cdf_exact_x")
        return dict(p=p, n=n, cdf=cdf, ws=ws, cap=cap, x=x)

    def search_x(self, h: dict, draws: Optional[torch.Tensor], m: int, out: torch.Tensor, systematic: bool = False,
                 u0: float = 0.0) -> torch.Tensor:
        """Local ancestor index of every (replicated) draw whose ancestor lives on this rank, -1 elsewhere."""
        ovf = self.ws.i32("cdfx_ovf", 1)
        _lib.check(self.lib.tb_cdf_search_x(ptr(h["p"]) if h["n"] else None, h["n"], ptr(h["cdf"]), ptr(h["ws"]),
                                            h["cap"], C.byref(h["x"]), ptr(draws), int(m), int(systematic), float(u0),
                                            ptr(out), ptr(ovf), stream_ptr()), "tb_cdf_search_x")
        self._pending_status = getattr(self, "_pending_status", [])
        self._pending_status.append((h["ws"][:64].view(torch.int32), ovf if systematic else None))
        if not self.defer_checks:
            self.check_pending()
        return out

    defer_checks = False      # the sampler sets it: status words are then read once per iteration (check_pending)

    def check_pending(self) -> None:
        """Read the status words of the sharded cdf calls since the last check (one host round trip for all)."""
        pend, self._pending_status = getattr(self, "_pending_status", []), []
        for status, ovf in pend:
            st = status.cpu().numpy()                                 # {tiles, segments, runs, hard tiles, error}
            self.last_cdf_status = st
            if st[4] != 0:
                raise RuntimeError(f"sharded exact cdf failed with code {int(st[4])} (3: a peer GPU did not answer, "
                                   f"4/5: table capacity, 6: refuted binade hypothesis); tiles={int(st[0])}, "
                                   f"hard={int(st[3])}")
            if ovf is not None and int(ovf.item()):
                raise IndexError("systematic resampling walked past the last weight (tools.py:223-225)")
        if self.comm.fast is not None:
            self.comm.fast.check()

    def peer_rows(self, per: int, d: int) -> "PeerRows":
        """Active-set buffers in peer-mapped memory for `per` walker slots per rank (pooled like the other symmetric
        allocations: the views a Sampler holds into them stay its own until it dies)."""
        rows = getattr(self, "_peer_rows", None)
        if rows is None or rows.per != per or rows.d != d:
            key = ("rows", per, d) + self._pool_tag
            rows = _pool_acquire(key, lambda: PeerRows(self.lib, self.device, self.comm, per, d))
            self._leased.append((key, rows))
            self._peer_rows = rows
        return rows

    def prepare(self, per: int, d: int, n_global_hint: int) -> None:
        """Set up the lazily created symmetric allocations now (Sampler construction) instead of inside the first
        iteration that trims / resamples."""
        if self.comm.fast is not None:
            self.peer_rows(per, d)
        if self.xgpu is not None:
            self._cdf_tables(int(self.lib.tb_cdf_tile_cap(max(int(n_global_hint), 1), 64 * self.comm.world)))

    def sharded_search(self, p: torch.Tensor, n: int, seg_begin: torch.Tensor, draws: torch.Tensor,
                       out: torch.Tensor, name: str, n_global: Optional[int] = None):
        """Ancestor index of every (replicated) multinomial draw whose ancestor lives on this rank, -1 elsewhere;
        identical to numpy's ``choice`` on the global weight vector for any number of GPUs."""
        if n_global is None:
            n_global = self.g_int(n)
        h = self.cdf_x(p, n, seg_begin, n_global, name)
        self.search_x(h, draws, int(draws.numel()), out)
        return out, None


def sharded_mode_moments(core, idx, counts, n_trim_local: int, m_total: int, cmean, scatter):
    """Count-weighted mean / scatter of the global resampled multiset (student.py:63)."""
    k, ens, d = core.k, core.ensemble, core.ensemble.n_dim
    lib, st = k.lib, stream_ptr()
    mws = k.ws.bytes("mom", lib.tb_moments_workspace_bytes(d))
    if n_trim_local:
        _lib.check(lib.tb_moments_partial(ptr(ens.u), ptr(idx), None, ptr(counts), n_trim_local, d, 1.0 / m_total, 1, 0,
                                          ptr(mws), ptr(cmean), None, st), "tb_moments_partial")
    else:
        cmean.zero_()
    k.comm.allreduce_sum_(cmean)
    if n_trim_local:
        _lib.check(lib.tb_moments_partial(ptr(ens.u), ptr(idx), None, ptr(counts), n_trim_local, d, 1.0 / m_total, 0, 1,
                                          ptr(mws), ptr(cmean), ptr(scatter), st), "tb_moments_partial")
    else:
        scatter.zero_()
    k.comm.allreduce_sum_(scatter)


class PeerRows:
    """Active-set buffers in peer-mapped memory: resampled rows are stored straight into their slot owner's buffer."""

    def __init__(self, lib, device, comm: Comm, per: int, d: int):
        import torch.distributed._symmetric_memory as symm

        self.lib, self.per, self.d = lib, per, d
        n = int(lib.tb_xrows_buffer_bytes(per, d)) // 8
        self.buf = symm.empty(n, dtype=F64, device=device)
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, torch.distributed.group.WORLD if comm.group is None else comm.group)
        torch.cuda.synchronize()
        torch.distributed.all_reduce(torch.zeros(1, device=device))
        self.ticket = torch.zeros(4, dtype=torch.int32, device=device)
        self.x = _lib.TbXcoll()
        self.x.rank, self.x.world, self.x.seq, self.x.cap_bytes = comm.rank, comm.world, 0, 0
        self.x.ticket = self.ticket.data_ptr()
        for r in range(comm.world):
            self.x.peer[r] = int(self.handle.buffer_ptrs[r])


def sharded_resample(core, weights: torch.Tensor, draws: Optional[torch.Tensor], systematic: bool = False,
                     u0: float = 0.0, k=None):
    """N global draws (multinomial: replicated uniforms; systematic: one uniform); returns this rank's block of
    resampled (u, logl) rows.  Every rank searches all N draws in the global exact cdf and stores the rows whose
    ancestors it holds straight into the active-set buffer of the rank that owns the walker slot (NVLink peer
    stores, no host synchronisation); without peer memory the rows travel by an all-to-all."""
    k = core.k if k is None else k            # (the side-stream kernels object when the resampling overlaps the Trainer)
    ens, comm = core.ensemble, k.comm
    n_glob = core.n_global
    d = ens.n_dim
    idx = k.ws.i64("res_idx", n_glob)
    h = k.cdf_x(weights, ens.n_total, core.generation_bounds(), ens.n_total_global, "cdf")
    k.search_x(h, draws, n_glob, idx, systematic=systematic, u0=u0)
    core.trace["resample_idx"] = idx
    if comm.fast is not None:
        rows = k.peer_rows(core.n_local, d)
        rows.x.seq += 1
        _lib.check(k.lib.tb_xrows_scatter(ptr(ens.u), ptr(ens.logl), d, ptr(idx), n_glob, core.n_local, C.byref(rows.x),
                                          ptr(comm.fast.err), stream_ptr()), "tb_xrows_scatter")
        off = int(k.lib.tb_xrows_offset(core.n_local, d, rows.x.seq & 1))
        u = rows.buf[off: off + core.n_local * d].view(core.n_local, d)
        logl = rows.buf[off + core.n_local * d: off + core.n_local * (d + 1)]
        return u, logl
    own = torch.nonzero(idx >= 0).flatten()
    rows = torch.empty((own.numel(), d + 1), dtype=F64, device=core.device)
    if own.numel():
        src = idx[own].contiguous()
        u = torch.empty((own.numel(), d), dtype=F64, device=core.device)
        l = torch.empty(own.numel(), dtype=F64, device=core.device)
        _lib.check(k.lib.tb_gather_rows(ptr(ens.u), ptr(ens.logl), d, ptr(src), own.numel(), ptr(u), ptr(l),
                                        stream_ptr()), "tb_gather_rows")
        rows[:, :d] = u
        rows[:, d] = l
    lo, hi = core.slot_offset, core.slot_offset + core.n_local
    mine = exchange_owned_rows(comm, rows, own, n_glob, lo, hi)
    return mine[:, :d].contiguous(), mine[:, d].contiguous()


def sharded_mcmc_loop(core, params, tape_ref, u, logl, qcur, ws, ctrl, n_min: int, n_cap: int, chunk: int):
    """One Metropolis step per launch; the per-step totals (sum alpha per mode, accepted, proposals,
    error) are all-reduced before sigma adaptation and the stop rule (tb_mcmc_update)."""
    k, comm = core.k, core.comm
    lib, sp = k.lib, stream_ptr()
    K = params.n_modes
    tot = ctrl[8 + 3 * K: 8 + 4 * K + 3]
    launched = 0
    next_check = n_min
    while True:
        _lib.check(lib.tb_mcmc_steps(core.n_local, C.byref(params), tape_ref, None, ptr(u), ptr(logl), ptr(qcur),
                                     ptr(ws), ptr(ctrl), 1, sp), "tb_mcmc_steps")
        comm.allreduce_sum_(tot)
        _lib.check(lib.tb_mcmc_update(C.byref(params), ptr(ctrl), sp), "tb_mcmc_update")
        launched += 1
        if launched >= next_check or launched >= n_cap:
            h = ctrl.cpu().numpy()
            if h[1] != 0.0 or launched >= n_cap:
                return h, launched
            next_check = launched + chunk
