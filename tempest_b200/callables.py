"""User callables on the generic path (reference: ``SamplerCore._log_like``, tempest/core.py:317-358;
``prior_transform`` per row, tempest/mcmc.py:157 and steps/mutate.py:103).

Registry objects (tempest_b200/registry.py) are evaluated inside the fused step kernel.  Anything
else is called here, between the proposal and the accept kernels of a split Metropolis step:

* numpy callables (what the reference accepts): the proposals make one device->host->device round
  trip per step.  ``log_likelihood`` receives the whole ``[N, D]`` batch when ``vectorize=True`` and one
  row at a time otherwise (core.py:321-326); ``prior_transform`` is called on the batch when a probe
  shows that the batched call equals the per-row calls, else row by row like the reference;
* device callables, marked with :func:`device_callable`: they receive and return CUDA fp64 torch
  tensors (``[N, D] -> [N, D]`` and ``[N, D] -> [N]``) and nothing leaves the GPU.
"""
from __future__ import annotations

from typing import Callable

import numpy as np
import torch

from .registry import is_registry_likelihood, is_registry_prior

F64 = torch.float64


def device_callable(fn: Callable) -> Callable:
    """Mark ``fn`` as operating on CUDA fp64 torch tensors (batched)."""
    fn._tb_device = True
    return fn


def _is_device(fn) -> bool:
    return bool(getattr(fn, "_tb_device", False))


class CallableBridge:
    def __init__(self, config, device: torch.device):
        self.device = device
        self.n_dim = config.n_dim
        self.vectorize = bool(config.vectorize)
        self.prior_fn = config.prior_transform
        wrapped = config.log_likelihood
        self.like_fn = wrapped
        inner = wrapped.f if hasattr(wrapped, "f") else wrapped
        self.like_inner = inner
        self.prior_registry = is_registry_prior(self.prior_fn)
        self.like_registry = is_registry_likelihood(inner) and not getattr(wrapped, "args", None) \
            and not getattr(wrapped, "kwargs", None)
        self.prior_device = _is_device(self.prior_fn)
        self.like_device = _is_device(inner)
        self.external = not (self.prior_registry and self.like_registry)
        self.prior_batched = True
        if not self.prior_registry and not self.prior_device:
            self.prior_batched = self._probe_prior_batched()
        self.n_like_calls = 0

    def _probe_prior_batched(self) -> bool:
        d = self.n_dim
        probe = (np.arange(3 * d, dtype=float).reshape(3, d) + 0.5) / (3 * d)
        rows = np.array([np.asarray(self.prior_fn(probe[i].copy()), dtype=float) for i in range(3)])
        try:
            batch = np.asarray(self.prior_fn(probe.copy()), dtype=float)
        except Exception:
            return False
        return batch.shape == rows.shape and np.array_equal(batch, rows)

    # x = prior_transform(u)
    def prior(self, u: torch.Tensor, core) -> torch.Tensor:
        if self.prior_registry:
            return core.registry_transform(u)
        if self.prior_device:
            x = self.prior_fn(u)
            return x.to(F64).contiguous()
        h = u.detach().cpu().numpy()
        if self.prior_batched:
            x = np.asarray(self.prior_fn(h), dtype=float)
        else:
            x = np.array([self.prior_fn(h[i]) for i in range(h.shape[0])], dtype=float)   # mcmc.py:157
        return torch.as_tensor(np.ascontiguousarray(x), dtype=F64).to(self.device)

    # logl = log_likelihood(x)
    def like(self, x: torch.Tensor) -> torch.Tensor:
        self.n_like_calls += 1
        if self.like_device:
            out = self.like_fn(x)
            return out.to(F64).reshape(-1).contiguous()
        h = x.detach().cpu().numpy()
        if self.vectorize:
            out = np.asarray(self.like_fn(h), dtype=float)                                # core.py:321-322
        else:
            out = np.array([self.like_fn(h[i]) for i in range(h.shape[0])], dtype=float)  # core.py:323-326
        if out.shape != (h.shape[0],):
            raise ValueError(f"log_likelihood returned shape {out.shape}, expected ({h.shape[0]},)")
        return torch.as_tensor(np.ascontiguousarray(out), dtype=F64).to(self.device)
