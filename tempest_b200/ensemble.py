"""Device-resident persistent ensemble and the ``StateManager``-shaped view over it.

Replaces tempest/state_manager.py (list-of-arrays history, re-concatenated on every access,
:313-314; defensive copies on every get/set, :668-674) with pre-allocated SoA tensors in HBM:

    u[cap, D] row-major fp64    unit-cube coordinates of every particle ever drawn
    logl[cap] fp64              their log-likelihoods
    C[cap]    fp64              cached log-mixture  C_s = LSE_t(log n_t + beta_t l_s - logZ_t)
    gen_beta/gen_logz/gen_logn  per-generation scalars (host lists + device mirror)

``x`` is not stored: registry priors are elementwise functions of ``u`` and are recomputed
on demand by ``tb_transform`` (halves history memory and gather traffic).
Capacity grows geometrically; appends are device-to-device copies plus ``tb_mixture_append``.
PyTorch is used for allocation, streams and host copies only.
"""
from __future__ import annotations

import ctypes as C
import weakref
import math
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib

CURRENT_STATE_KEYS = frozenset({
    "u", "x", "logl", "assignments", "blobs", "acceptance", "steps", "efficiency", "ess", "cv",
    "beta", "logz", "calls", "iter"})  # state_manager.py:7-24
HISTORY_STATE_KEYS = frozenset({
    "u", "x", "logl", "blobs", "iter", "logz", "calls", "steps", "efficiency", "ess", "cv",
    "acceptance", "beta"})  # state_manager.py:26-42


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr() -> C.c_void_p:
    """cudaStream_t of torch's current stream on the current device (every C-ABI call takes it).  The raw accessor is
    ~50x cheaper than building a torch.cuda.Stream object -- at ~2 000 calls per run that was 6 % of the host path."""
    if _raw_stream is not None:
        return C.c_void_p(_raw_stream(torch.cuda.current_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


class PersistentEnsemble:
    """All particles ever drawn, resident in HBM."""

    def __init__(self, n_dim: int, device: torch.device, capacity: int = 0, world: int = 1):
        self.n_dim = int(n_dim)
        self.device = device
        self.world = int(world)          # shards of equal size: global counts = local counts * world
        self.gen_n_local: List[int] = []
        self.lib = _lib.load()
        self.n_total = 0
        self.cap = 0
        self.u = torch.empty((0, self.n_dim), dtype=torch.float64, device=device)
        self.logl = torch.empty((0,), dtype=torch.float64, device=device)
        self.C = torch.empty((0,), dtype=torch.float64, device=device)
        self.gen_beta: List[float] = []
        self.gen_logz: List[float] = []
        self.gen_n: List[int] = []
        self._gcap = 0
        self._dev_gens = torch.empty((3, 0), dtype=torch.float64, device=device)
        if capacity:
            self._reserve(capacity)

    # -- capacity -----------------------------------------------------------------------
    RESERVE_GENERATIONS = 48      # a C4-like run stores ~36 generations; growth beyond this doubles

    def _reserve(self, need: int) -> None:
        if need <= self.cap:
            return
        # first allocation: room for RESERVE_GENERATIONS generations of this size, so that neither the
        # store nor the per-N_total scratch buffers are re-allocated (cudaMalloc/cudaFree stalls) mid-run
        first = need * self.RESERVE_GENERATIONS if self.cap == 0 else 0
        new_cap = max(need, int(self.cap * 2), first, 1024)
        for name, shape in (("u", (new_cap, self.n_dim)), ("logl", (new_cap,)), ("C", (new_cap,))):
            old = getattr(self, name)
            new = torch.empty(shape, dtype=torch.float64, device=self.device)
            if self.n_total:
                new[: self.n_total].copy_(old[: self.n_total])
            setattr(self, name, new)
        self.cap = new_cap

    def _sync_gens(self) -> None:
        T = len(self.gen_beta)
        if T > self._gcap:
            self._gcap = max(64, 2 * T)
            self._dev_gens = torch.zeros((3, self._gcap), dtype=torch.float64, device=self.device)
        host = torch.from_numpy(np.stack([np.asarray(self.gen_beta, dtype=float),
                                          np.asarray(self.gen_logz, dtype=float),
                                          np.log(np.asarray(self.gen_n))]))   # np.log as state_manager.py:468
        self._dev_gens[:, :T].copy_(host)

    @property
    def T(self) -> int:
        return len(self.gen_beta)

    @property
    def n_total_global(self) -> int:
        return self.n_total * self.world

    # -- append one generation (commit, state_manager.py:356-416) -------------------------
    def append(self, u_new: torch.Tensor, logl_new: torch.Tensor, beta: float, logz: float) -> None:
        n_new = int(logl_new.shape[0])
        n_old = self.n_total
        self._reserve(n_old + n_new)
        self.u[n_old:n_old + n_new].copy_(u_new)
        self.logl[n_old:n_old + n_new].copy_(logl_new)
        self.gen_beta.append(float(beta))
        self.gen_logz.append(float(logz))
        self.gen_n.append(n_new * self.world)      # the mixture weights n_t / N use GLOBAL counts
        self.gen_n_local.append(n_new)
        self._sync_gens()
        g = self._dev_gens
        _lib.check(self.lib.tb_mixture_append(
            ptr(self.logl), ptr(self.C), n_old, n_new, ptr(g[0]), ptr(g[1]), ptr(g[2]), self.T,
            stream_ptr()), "tb_mixture_append")
        self.n_total = n_old + n_new

    def rebuild_mixture(self) -> None:
        """Full N_total x T rebuild (used by tests to check the incremental path)."""
        g = self._dev_gens
        _lib.check(self.lib.tb_mixture_build(
            ptr(self.logl), ptr(self.C), self.n_total, ptr(g[0]), ptr(g[1]), ptr(g[2]), self.T,
            stream_ptr()), "tb_mixture_build")

    def all_warmup(self) -> bool:
        """True while every stored generation was drawn at beta = 0."""
        return all(b == 0.0 for b in self.gen_beta)


class DeviceState:
    """``StateManager`` surface used by the sampler and by the reference's tests
    (get_current / get_history / get_history_length / get_last_history /
    compute_logw_and_logz), backed by the device ensemble.  Arrays are handed out as numpy
    copies (the reference copies on every get, state_manager.py:668-674)."""

    def __init__(self, n_dim: int, core=None):
        self.n_dim = n_dim
        self._core = weakref.proxy(core) if core is not None else None
        self._current: Dict[str, object] = {k: None for k in CURRENT_STATE_KEYS}
        self._history: Dict[str, list] = {k: [] for k in HISTORY_STATE_KEYS if k not in ("u", "x", "logl", "blobs")}
        self._history["blobs"] = []

    # -- current ------------------------------------------------------------------------
    def _export(self, key, value):
        if isinstance(value, torch.Tensor):
            return value.detach().cpu().numpy().copy()
        if isinstance(value, np.ndarray):
            return value.copy()
        return value

    def get_current(self, key: Optional[str] = None):
        if key is None:
            return {k: self.get_current(k) for k in CURRENT_STATE_KEYS}
        if key not in CURRENT_STATE_KEYS:
            raise KeyError(f"Invalid current state key: '{key}'")
        if key == "x" and self._current["x"] is None and self._current["u"] is not None and self._core is not None:
            return self._core.transform_to_x(self._current["u"]).cpu().numpy()
        return self._export(key, self._current[key])

    def set_current(self, key: str, value) -> None:
        if key not in CURRENT_STATE_KEYS:
            raise KeyError(f"Invalid current state key: '{key}'")
        self._current[key] = value

    def update_current(self, data: dict) -> None:
        for k, v in data.items():
            self.set_current(k, v)

    def raw(self, key: str):
        return self._current[key]

    # -- history ------------------------------------------------------------------------
    def get_history_length(self) -> int:
        return len(self._history["beta"])

    def commit_scalars(self) -> None:
        for k in self._history:
            if k == "blobs":
                continue
            v = self._current[k]
            if v is not None:
                self._history[k].append(v)

    def _particle_history(self, key: str) -> List[np.ndarray]:
        ens: PersistentEnsemble = self._core.ensemble
        if ens.n_total == 0:
            return []
        if key == "logl":
            flat = ens.logl[: ens.n_total].cpu().numpy()
        elif key == "u":
            flat = ens.u[: ens.n_total].cpu().numpy()
        else:
            flat = self._core.transform_to_x(ens.u[: ens.n_total]).cpu().numpy()
        out, o = [], 0
        for n in ens.gen_n_local:
            out.append(flat[o:o + n].copy())
            o += n
        return out

    def get_history(self, key: str, index: Optional[int] = None, flat: bool = False):
        if key not in HISTORY_STATE_KEYS:
            raise KeyError(f"Invalid history state key: '{key}'")
        items = self._particle_history(key) if key in ("u", "x", "logl") else list(self._history[key])
        if index is None:
            if flat:
                return np.concatenate(items)
            return np.array(items)
        if index >= len(items) or index < 0:
            raise IndexError(f"Index {index} out of range for history key '{key}'")
        return items[index]

    def get_last_history(self, key: str, default=None):
        if key not in HISTORY_STATE_KEYS:
            raise KeyError(f"Invalid history state key: '{key}'")
        if self.get_history_length() == 0:
            return default
        return self.get_history(key, self.get_history_length() - 1)

    def compute_logw_and_logz(self, beta_final: float = 1.0, normalize: bool = True):
        """state_manager.py:418-480 on the device; returns (numpy logw, logz)."""
        if self.get_history_length() == 0:
            return np.array([]), -np.inf
        return self._core.logw_and_logz(beta_final, normalize)

    def compute_results(self) -> dict:
        out = {k: self.get_history(k) for k in HISTORY_STATE_KEYS if k != "blobs"}
        out["logw"], _ = self.compute_logw_and_logz(1.0)
        return out
