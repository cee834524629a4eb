"""Sources of random variates for the device path.

``PhiloxSource`` (production): counter-based Philox4x32-10 keyed by (seed, PS iteration) with
counters (walker slot, MCMC step, purpose, sub-draw); variates are generated inside the kernels,
so results do not depend on grid shape or on how walkers are sharded over GPUs.

``TapeSource`` (parity): replays variates recorded from the reference's own legacy MT19937
stream: per PS iteration a dict with ``prior_u [N,D]``, ``train_u [4 n_trim]``,
``resample_u [N] | [1]``, and for the MCMC call ``gamma[step][k]`` (standard-gamma variates),
``z[step][k] -> [attempts, D]`` normals and ``acc_u[step][k]`` uniforms.  The reference draws
all of these from one sequential global stream with data-dependent counts, so bit-level parity
is only definable on recorded tapes.
"""
from __future__ import annotations

import ctypes as C
import secrets
from typing import List, Optional

import numpy as np
import torch

from . import _lib
from .ensemble import ptr, stream_ptr

PURPOSE_RESAMPLE = 4
PURPOSE_TRAIN = 5


class PhiloxSource:
    mode = 0  # TB_RNG_PHILOX

    def __init__(self, seed: Optional[int], device: torch.device):
        self.seed = int(seed) if seed is not None else secrets.randbits(62)
        self.device = device
        self.iteration = 0
        self.epoch = 0
        self.lib = _lib.load()

    @property
    def key(self) -> int:
        """Philox key material: the seed, advanced per run epoch.  A second run() on the same Sampler continues on top
        of the stored history with iter reset to 0 (core.py:376-381); keyed by (seed, iteration) alone it would redraw
        the first run's particles bit for bit (the reference's MT19937 stream simply continues)."""
        return (self.seed + self.epoch * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF

    def set_epoch(self, epoch: int) -> None:
        self.epoch = int(epoch)

    def begin_iteration(self, iteration: int) -> None:
        self.iteration = int(iteration)

    def _uniform(self, purpose: int, n: int, offset: int = 0) -> torch.Tensor:
        out = torch.empty(n, dtype=torch.float64, device=self.device)
        _lib.check(self.lib.tb_philox_uniform(self.key, self.iteration, purpose, offset, n, ptr(out),
                                              stream_ptr()), "tb_philox_uniform")
        return out

    def prior_u(self, n: int, d: int) -> Optional[torch.Tensor]:
        return None  # generated in-kernel by tb_prior_draw

    def train_u(self, m: int) -> torch.Tensor:
        return self._uniform(PURPOSE_TRAIN, m)

    def resample_u(self, n: int) -> torch.Tensor:
        return self._uniform(PURPOSE_RESAMPLE, n)

    def resample_u0(self) -> float:
        return float(self._uniform(PURPOSE_RESAMPLE, 1, offset=1 << 40).item())

    def inf_pick(self, pool: torch.Tensor, size: int) -> torch.Tensor:
        u = self._uniform(PURPOSE_RESAMPLE, size, offset=1 << 41)
        return pool[(u * pool.numel()).long().clamp_(max=pool.numel() - 1)]

    def mcmc_tape(self, n: int, d: int):
        return None


class TapeSource:
    mode = 1  # TB_RNG_TAPE

    def __init__(self, tapes: List[dict], device: torch.device):
        self.tapes = tapes
        self.device = device
        self.cursor = -1
        self.seed = 0
        self.key = 0
        self.iteration = 0
        self._keep = []

    def set_epoch(self, epoch: int) -> None:
        pass

    def begin_iteration(self, iteration: int) -> None:
        self.cursor += 1
        self.iteration = int(iteration)
        self._keep = []
        if self.cursor >= len(self.tapes):
            raise RuntimeError("tape exhausted: the device path ran more PS iterations than were recorded")

    @property
    def tape(self) -> dict:
        return self.tapes[self.cursor]

    def _dev(self, a, dtype=torch.float64) -> torch.Tensor:
        t = torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(self.device)
        self._keep.append(t)
        return t

    def _need(self, key: str):
        if key not in self.tape:
            raise RuntimeError(f"tape of iteration {self.cursor} has no '{key}': control flow diverged from the recording")
        return self.tape[key]

    def prior_u(self, n: int, d: int) -> torch.Tensor:
        a = np.asarray(self._need("prior_u"))
        assert a.shape == (n, d)
        return self._dev(a)

    def train_u(self, m: int) -> torch.Tensor:
        a = np.asarray(self._need("train_u"))
        if a.shape[0] != m:
            raise RuntimeError(f"tape holds {a.shape[0]} training uniforms, device path needs {m} (trim set differs)")
        return self._dev(a)

    def resample_u(self, n: int) -> torch.Tensor:
        a = np.asarray(self._need("resample_u"))
        assert a.shape[0] == n
        return self._dev(a)

    def resample_u0(self) -> float:
        return float(np.asarray(self._need("resample_u"))[0])

    def inf_pick(self, pool: torch.Tensor, size: int) -> torch.Tensor:
        return self._dev(np.asarray(self._need("inf_pick")), dtype=torch.int64)

    def mcmc_tape(self, n: int, d: int):
        """Flatten the recorded MCMC variates into the tb_tape layout."""
        t = self.tape
        steps = len(t["acc_u"])
        acc_u = self._dev(np.stack(t["acc_u"]))
        have_gamma = len(t.get("gamma", [])) == steps
        gamma = self._dev(np.stack(t["gamma"])) if have_gamma else None
        cnt = np.empty((steps, n), dtype=np.int32)
        chunks = []
        for s in range(steps):
            for k in range(n):
                zk = np.asarray(t["z"][s][k]).reshape(-1, d)
                cnt[s, k] = zk.shape[0]
                chunks.append(zk.ravel())
        off = np.zeros(steps * n, dtype=np.int64)
        np.cumsum(cnt.ravel()[:-1].astype(np.int64) * d, out=off[1:])
        z = self._dev(np.concatenate(chunks))
        z_off = self._dev(off, dtype=torch.int64)
        z_cnt = self._dev(cnt, dtype=torch.int32)
        tape = _lib.TbTape(gamma=ptr(gamma).value, acc_u=ptr(acc_u).value, z=ptr(z).value,
                           z_off=ptr(z_off).value, z_cnt=ptr(z_cnt).value, steps=steps)
        return tape
