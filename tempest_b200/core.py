"""Orchestrator of the device Persistent Sampling loop (reference: tempest/core.py).

``execute_iteration`` = reweight -> train -> resample -> mutate -> commit (core.py:162-185);
``run_sampling`` = the ``while _not_termination()`` loop (core.py:110-160, 360-374);
``compute_posterior`` / ``compute_evidence`` (core.py:187-247).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import Optional, Union

import numpy as np
import torch

from . import _lib
from .config import SamplerConfig
from .dist import Comm, shard_bounds
from .ensemble import DeviceState, PersistentEnsemble, ptr, stream_ptr
from .registry import is_registry_likelihood, is_registry_prior
from .rng import PhiloxSource
from .steps import F64, Kernels, ModeStats, Mutator, Resampler, Reweighter, Trainer


class SamplerCore:
    def __init__(self, config: SamplerConfig):
        self.config = config
        self._check_supported(config)
        if not torch.cuda.is_available():
            raise RuntimeError("tempest_b200 needs a CUDA device (sm_100a): there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.lib = _lib.load()
        # particles shard over the ranks of an initialised torch.distributed job
        self.comm = Comm()
        self.n_global = config.n_particles
        if self.comm.on:
            from .sharded import ShardedKernels

            if config.random_state is None:
                raise ValueError("pass random_state on >1 GPU: every rank must draw the same replicated uniforms")
            lo, hi = shard_bounds(self.n_global, self.comm.world, self.comm.rank)
            self.slot_offset, self.n_local = lo, hi - lo
            self.k = ShardedKernels(self.device, self.comm)
            self.k.defer_checks = True
        else:
            self.slot_offset, self.n_local = 0, self.n_global
            self.k = Kernels(self.device)
        self.ensemble = PersistentEnsemble(config.n_dim, self.device, world=self.comm.world)
        self.state = DeviceState(config.n_dim, core=self)
        self.rng = PhiloxSource(config.random_state, self.device)
        self.trace: dict = {}
        self.kernel_timing = None       # dict: bench.py collects (CUDA events, work) per launch of the two hot kernels
        self.n_mcmc_launches = 0
        self.n_total = 0
        self.logz_err = None
        self._weights = None
        # registry objects run inside the fused step kernel; anything else goes through the split step
        # (tb_mcmc_propose -> callables -> tb_mcmc_accept), see callables.py
        from .callables import CallableBridge

        self.bridge = CallableBridge(config, self.device)
        if self.bridge.external and self.comm.on:
            raise NotImplementedError("arbitrary callables are single-GPU for now; registry priors/likelihoods shard")
        like = self.bridge.like_inner
        self._like = like
        self._like_id = like.kernel_id if self.bridge.like_registry else -1
        self._like_params = (torch.as_tensor(like.dparams(), dtype=F64).to(self.device)
                             if self.bridge.like_registry else None)
        self._prior_params = (torch.as_tensor(config.prior_transform.dparams(), dtype=F64).to(self.device)
                              if self.bridge.prior_registry else None)
        kinds = np.zeros(config.n_dim, dtype=np.uint8)
        for j in (config.periodic or []):
            kinds[j] = 1
        for j in (config.reflective or []):
            kinds[j] = 2
        self._bc = torch.as_tensor(kinds).to(self.device) if kinds.any() else None
        self.clusterer = None
        self.assign = None                                     # int32 device labels of the active walkers
        if config.clustering:                                  # core.py:54-69
            if self.comm.on:
                raise NotImplementedError("clustering=True is not sharded yet; run it on one GPU or pass "
                                          "clustering=False (SURVEY 8e)")
            from .cluster import HierarchicalGaussianMixture

            cap = config.n_max_clusters
            self.clusterer = HierarchicalGaussianMixture(
                self.k, n_init=1, max_iterations=1000 if cap is None else cap - 1,
                min_points=None if cap is None else 4 * config.n_dim,
                threshold_modifier=config.split_threshold, covariance_type="full", verbose=False,
                normalize=config.normalize)
        # independent branches of one iteration run on a side stream (single GPU, ESS mode, one mode,
        # multinomial resampling): the cv diagnostic and the resampling overlap with the Trainer
        self.overlap = (config.volume_variation is None and not config.clustering and config.resample == "mult"
                        and (not self.comm.on or self.k.comm.fast is not None))
        if os.environ.get("TEMPEST_B200_OVERLAP", "1") == "0":
            self.overlap = False
        self.side = torch.cuda.Stream(self.device) if self.overlap else None
        if self.overlap and self.comm.on:
            # sharded: the side branches issue collectives of their own, so they get their own peer-memory channel
            # (a second Comm / PeerCollectives / table memory); both streams issue the same call sequence on every rank
            from .dist import Comm as _Comm

            self.k_side = ShardedKernels(self.device, _Comm(), role="side")
            self.k_side.defer_checks = True
        else:
            self.k_side = Kernels(self.device, role="side") if self.overlap else None
        if self.comm.on:
            from .ensemble import PersistentEnsemble as _PE

            for kk in (self.k, self.k_side):
                if kk is not None:
                    kk.prepare(self.n_local, config.n_dim, self.n_global * _PE.RESERVE_GENERATIONS)
        self._cv_pending = None
        self._nvtx = os.environ.get("TEMPEST_B200_NVTX", "0") == "1"
        self._nvtx_open = False
        self.reweighter = Reweighter(self)
        self.trainer = Trainer(self)
        self.resampler = Resampler(self)
        self.mutator = Mutator(self)

    # -- what the CUDA path supports ----------------------------------------------------------
    @staticmethod
    def _check_supported(cfg: SamplerConfig) -> None:
        if cfg.blobs_dtype is not None:
            raise NotImplementedError("blobs are not carried by the device ensemble (state_manager.py blobs)")
        if cfg.pool is not None:
            raise NotImplementedError("pool is meaningless on the vectorised CUDA path")
        like = cfg.log_likelihood.f if hasattr(cfg.log_likelihood, "f") else cfg.log_likelihood
        for obj, ok in ((like, is_registry_likelihood(like)), (cfg.prior_transform, is_registry_prior(cfg.prior_transform))):
            if ok and obj.n_dim != cfg.n_dim:
                raise ValueError("registry prior / likelihood dimension does not match n_dim")

    # -- helpers used by the steps ---------------------------------------------------------------
    @property
    def warmup_regime(self) -> bool:
        return self.ensemble.all_warmup()

    def zero_assignments(self, n: int) -> np.ndarray:
        """The all-zero label vector of an unclustered iteration (resample.py:71; one shared read-only array instead of
        a fresh 8 MB fill per iteration)."""
        z = getattr(self, "_zero_assign", None)
        if z is None or z.shape[0] != n:
            z = np.zeros(n, dtype=int)
            z.flags.writeable = False
            self._zero_assign = z
        return z

    def generation_bounds(self) -> torch.Tensor:
        """Local start positions [T+1] of every stored generation in this rank's history shard."""
        key = (len(self.ensemble.gen_n_local), self.ensemble.n_total)
        cached = getattr(self, "_gen_bounds", None)
        if cached is None or cached[0] != key:          # (two callers per iteration; one host -> device copy)
            b = np.concatenate([[0], np.cumsum(self.ensemble.gen_n_local)]).astype(np.int64)
            cached = self._gen_bounds = (key, torch.as_tensor(b).to(self.device))
        return cached[1]

    def weights_buffer(self) -> torch.Tensor:
        n = self.ensemble.n_total
        self.k.ws.hint = self.ensemble.cap
        if self._weights is None or self._weights.numel() < n:
            self._weights = torch.empty(max(int(n * 2), self.ensemble.cap, 1024), dtype=F64, device=self.device)
        return self._weights[:n]

    def mcmc_params(self, beta: float, mode_stats: Optional[ModeStats]) -> _lib.TbMcmcParams:
        cfg = self.config
        p = _lib.TbMcmcParams()
        p.n_dim = cfg.n_dim
        p.n_modes = mode_stats.K if mode_stats is not None else 1
        p.sampler = 1 if cfg.sample == "rwm" else 0
        p.rng_mode = self.rng.mode
        p.like_id = self._like_id
        p.prior_id = cfg.prior_transform.kernel_id if self.bridge.prior_registry else -1
        p.n_steps = cfg.n_steps
        p.n_max = cfg.n_max_steps
        p.beta = float(beta)
        p.seed = self.rng.key
        p.iteration = int(self.rng.iteration)
        p.slot_offset = self.slot_offset
        p.n_global = self.n_global
        p.defer_update = 1 if self.comm.on else 0
        p.like_params = self._like_params.data_ptr() if self._like_params is not None else None
        p.prior_params = self._prior_params.data_ptr() if self._prior_params is not None else None
        if mode_stats is not None:
            p.mode_mean = mode_stats.means.data_ptr()
            p.mode_chol = mode_stats.chol_covariances.data_ptr()
            p.mode_inv = mode_stats.inv_covariances.data_ptr()
            p.mode_dof = mode_stats.degrees_of_freedom.data_ptr()
        p.bc_kind = self._bc.data_ptr() if self._bc is not None else None
        return p

    # -- cv diagnostic on the side stream (reweight.py:417-419 computes it inline) -----------------------
    def begin_cv(self, w: torch.Tensor) -> None:
        ens = self.ensemble
        n, d = ens.n_total, ens.n_dim
        ks = self.k_side
        ks.ws.hint = ens.cap
        w_cv = ks.ws.f64("cv_w", n)
        w_cv.copy_(w)                       # trim_weights renormalises w in place; cv uses the weights as they are now
        ready = torch.cuda.Event()
        ready.record()
        self.side.wait_event(ready)
        with torch.cuda.stream(self.side):
            if self.comm.on:
                ks.volume_variation_async(ens.u, w_cv, n, d)        # moments + all-reduces + Cholesky, all on the device
            else:
                ks.volume_variation_begin(ens.u, w_cv, n, d)
        self._cv_pending = (w_cv, n)

    def mid_cv(self) -> None:
        if self._cv_pending is not None and not self.comm.on:
            w_cv, n = self._cv_pending
            with torch.cuda.stream(self.side):
                self.k_side.volume_variation_mid(self.ensemble.u, w_cv, n, self.ensemble.n_dim)

    def end_cv(self) -> None:
        if getattr(self, "_cv_sharded", False):
            self._cv_sharded = False
            self.state.set_current("cv", self.k.volume_variation_result())
        if self._cv_pending is not None:
            self._cv_pending = None
            with torch.cuda.stream(self.side):
                cv = self.k_side.volume_variation_result() if self.comm.on else self.k_side.volume_variation_end()
            self.state.set_current("cv", cv)

    def transform_to_x(self, u: torch.Tensor) -> torch.Tensor:
        """x = prior_transform(u): tb_transform for registry priors, the user's callable otherwise."""
        if not self.bridge.prior_registry:
            return self.bridge.prior(u, self) if int(u.shape[0]) else torch.empty_like(u)
        return self.registry_transform(u)

    def registry_transform(self, u: torch.Tensor) -> torch.Tensor:
        n = int(u.shape[0])
        x = torch.empty_like(u)
        if n:
            p = self.mcmc_params(0.0, None)
            _lib.check(self.lib.tb_transform(ptr(u), n, C.byref(p), ptr(x), None, stream_ptr()), "tb_transform")
        return x

    def logw_and_logz(self, beta_final: float = 1.0, normalize: bool = True):
        ens = self.ensemble
        stats = self.k.probe(ens, beta_final, torch.zeros(16, dtype=F64, device=self.device))
        out = torch.empty(ens.n_total, dtype=F64, device=self.device)
        self.k.weights(ens, beta_final, stats, out, log=True)
        h = stats.cpu().numpy()
        logw = out.cpu().numpy()
        if not normalize:
            logw = logw + h[4] + np.log(ens.n_total_global)   # undo the normalisation: logw_s = a_s + log N
        return logw, float(h[4])

    # -- the PS loop ----------------------------------------------------------------------------
    def reset(self) -> None:
        """Forget the history but keep every device buffer (a fresh run on the same allocation; the
        reference needs a new Sampler for that because run() never clears its history, core.py:376-381)."""
        ens = self.ensemble
        ens.n_total = 0
        ens.gen_beta, ens.gen_logz, ens.gen_n, ens.gen_n_local = [], [], [], []
        st = self.state
        for key in st._history:
            st._history[key] = []
        for key in list(st._current):
            st._current[key] = None
        self.assign = None
        self._cv_pending = None
        self.trace = {}
        self._initialize_fresh()

    def _initialize_fresh(self) -> None:                 # core.py:376-381 (history is NOT cleared)
        self.state.update_current({"iter": 0, "calls": 0, "beta": 0.0, "logz": 0.0})
        # a run that starts on top of an existing history draws from a fresh Philox key (see PhiloxSource.key)
        self.run_epoch = getattr(self, "run_epoch", 0) + 1 if self.ensemble.T > 0 else 0
        self.rng.set_epoch(self.run_epoch)

    def _not_termination(self) -> bool:                  # core.py:360-374
        if self.ensemble.T == 0:
            return True
        h = self.k.probe(self.ensemble, 1.0).cpu().numpy()
        self._last_posterior_probe = h
        return 1.0 - float(self.state.raw("beta")) >= 1e-4 or h[3] < self.n_total

    def _stage(self, name: str):
        """CUDA-event stage timer (enabled by ``self.profile = True``; bench.py reads ``stage_ms``).  With
        TEMPEST_B200_NVTX=1 every stage is also an NVTX range (readable nsys / ncu timelines), at no cost otherwise."""
        if self._nvtx:
            if self._nvtx_open:
                torch.cuda.nvtx.range_pop()
            self._nvtx_open = name != "end"
            if self._nvtx_open:
                torch.cuda.nvtx.range_push("ps:" + name)
        if not getattr(self, "profile", False):
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self._stage_events.append((name, ev))

    def _flush_stages(self) -> None:
        if not getattr(self, "profile", False) or len(self._stage_events) < 2:
            return
        torch.cuda.synchronize()
        for (name, a), (_, b) in zip(self._stage_events[:-1], self._stage_events[1:]):
            self.stage_ms[name] = self.stage_ms.get(name, 0.0) + a.elapsed_time(b)
        self._stage_events = []

    def execute_iteration(self, save_every=None, t0: int = 0, export: bool = True) -> dict:
        if save_every is not None:                           # core.py:164-172
            it = int(self.state.raw("iter") or 0)
            if (it - t0) % int(save_every) == 0 and it != t0:
                self.save_sampler_state(self.config.output_dir / f"{self.config.output_label}_{it}.state")
        self.trace = {}
        if getattr(self, "profile", False):
            self._stage_events = []
            if not hasattr(self, "stage_ms"):
                self.stage_ms = {}
        self.rng.begin_iteration(int(self.state.raw("iter") or 0) + 1)
        self._stage("reweight")
        weights = self.reweighter.run()
        self._stage("train")
        mode_stats = self.trainer.run(weights)
        self.mid_cv()
        self._stage("resample")
        self.resampler.run(weights)
        self._stage("mutate")
        if self.comm.on and self.overlap:
            # the persistent kernels of the mutation / the next ESS search occupy every SM and wait for the peers inside
            # the kernel: all side-stream collectives must have completed on this GPU before they start
            torch.cuda.current_stream().wait_stream(self.side)
        self.mutator.run(mode_stats)
        self.end_cv()
        if self.comm.on:
            self.k.check_pending()          # status words of this iteration's sharded cdf calls / peer collectives
            if self.k_side is not None and hasattr(self.k_side, "check_pending"):
                self.k_side.check_pending()
        self._stage("commit")
        # commit (state_manager.py:356-416): particles to the device ensemble, scalars to host lists
        st = self.state
        self.ensemble.append(st.raw("u"), st.raw("logl"), float(st.raw("beta")), float(st.raw("logz")))
        st.commit_scalars()
        self._stage("end")
        self._flush_stages()
        self.last_mode_stats = mode_stats
        return st.get_current() if export else None

    def run_sampling(self, n_total: int = 4096, progress: bool = True, resume_state_path=None,
                     save_every: Optional[int] = None) -> None:
        if resume_state_path is not None:                   # core.py:118-126
            self.load_sampler_state(resume_state_path)
            t0 = int(self.state.raw("iter") or 0)
        else:
            t0 = 0
            self._initialize_fresh()
        self.t0 = t0
        self.n_total = int(n_total)
        pbar = None
        if progress:
            from tqdm import tqdm

            pbar = tqdm(desc="Iter", initial=t0)
        while self._not_termination():
            self.execute_iteration(save_every=save_every, t0=t0, export=False)
            if pbar is not None:
                st = self.state
                pbar.update(1)
                pbar.set_postfix(beta=st.raw("beta"), calls=st.raw("calls"), ESS=int(st.raw("ess")),
                                 logZ=st.raw("logz"), acc=st.raw("acceptance"), steps=st.raw("steps"))
        self.state.set_current("logz", float(self._last_posterior_probe[4]))   # core.py:149-150
        self.logz_err = None
        if save_every is not None:                          # core.py:153-156
            self.save_sampler_state(self.config.output_dir / f"{self.config.output_label}_final.state")
        if pbar is not None:
            pbar.close()

    # -- checkpoint / resume (core.py:249-315).  The reference pickles the whole sampler with dill and
    # its load path is broken upstream; this is a plain .npz of the persistent ensemble,
    # the per-generation scalars and the RNG position, written atomically. ---------------------------
    STATE_FORMAT = 1

    def save_sampler_state(self, path) -> None:
        from pathlib import Path

        path = Path(path)
        path.parent.mkdir(parents=True, exist_ok=True)
        if self.comm.on:
            # sharded run: every rank writes its own shard next to a manifest written by rank 0
            if self.comm.rank == 0:
                tmp = path.with_suffix(path.suffix + ".temp")
                with open(tmp, "wb") as f:
                    np.savez(f, format=np.array(self.STATE_FORMAT), world=np.array(self.comm.world),
                             n_dim=np.array(self.config.n_dim), n_particles=np.array(self.config.n_particles))
                tmp.replace(path)
            path = path.with_suffix(path.suffix + f".rank{self.comm.rank}")
        ens, st = self.ensemble, self.state
        n = ens.n_total
        out = dict(
            format=np.array(self.STATE_FORMAT), n_dim=np.array(self.config.n_dim),
            n_particles=np.array(self.config.n_particles),
            u=ens.u[:n].cpu().numpy(), logl=ens.logl[:n].cpu().numpy(),
            gen_beta=np.array(ens.gen_beta, dtype=float), gen_logz=np.array(ens.gen_logz, dtype=float),
            gen_n_local=np.array(ens.gen_n_local, dtype=np.int64),
            random_state=np.array(-1 if self.config.random_state is None else self.config.random_state),
            rng_seed=np.array(self.rng.seed, dtype=np.uint64), n_total=np.array(self.n_total),
        )
        for key, vals in st._history.items():
            if key != "blobs":
                out["hist_" + key] = np.array(vals, dtype=float)
        for key in ("iter", "calls", "beta", "logz", "steps", "acceptance", "efficiency", "ess", "cv"):
            v = st.raw(key)
            out["cur_" + key] = np.array(np.nan if v is None else float(v))
        for key in ("u", "logl", "assignments"):
            v = st.raw(key)
            if v is not None:
                out["cur_" + key] = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
        if self.clusterer is not None:                       # the fitted hierarchy: predict() runs on iterations
            out.update(self.clusterer.state_dict())          # with iter % cluster_every != 0 without a refit
        if self.assign is not None:
            out["core_assign"] = self.assign.detach().cpu().numpy()
        out["run_epoch"] = np.array(getattr(self, "run_epoch", 0))
        tmp = path.with_suffix(path.suffix + ".temp")
        with open(tmp, "wb") as f:
            np.savez(f, **out)
        tmp.replace(path)

    def load_sampler_state(self, path) -> None:
        from pathlib import Path

        path = Path(path)
        if self.comm.on:
            with np.load(path) as z:
                man = {k: z[k] for k in z.files}
            if "world" not in man or int(man["world"]) != self.comm.world:
                raise ValueError(f"state was written by {int(man.get('world', 1))} rank(s), this job has {self.comm.world}")
            path = path.with_suffix(path.suffix + f".rank{self.comm.rank}")
        with np.load(path) as z:
            d = {k: z[k] for k in z.files}
        if "world" in d and "u" not in d:
            raise ValueError(f"state was written by a sharded run of {int(d['world'])} ranks; load it under torchrun")
        if int(d["format"]) != self.STATE_FORMAT:
            raise ValueError(f"unknown state format {int(d['format'])}")
        if int(d["n_dim"]) != self.config.n_dim:
            raise ValueError(f"state has n_dim={int(d['n_dim'])}, sampler has n_dim={self.config.n_dim}")
        if "n_particles" in d and int(d["n_particles"]) != self.config.n_particles:
            raise ValueError(f"state has n_particles={int(d['n_particles'])}, sampler has "
                             f"n_particles={self.config.n_particles}")
        ens = PersistentEnsemble(self.config.n_dim, self.device, world=self.comm.world)
        n = int(d["logl"].shape[0])
        if n:
            ens.RESERVE_GENERATIONS = 1                     # size for the history plus room to continue
            ens._reserve(n + PersistentEnsemble.RESERVE_GENERATIONS * self.n_local)
            del ens.RESERVE_GENERATIONS
            ens.u[:n].copy_(torch.as_tensor(d["u"]).to(self.device))
            ens.logl[:n].copy_(torch.as_tensor(d["logl"]).to(self.device))
            ens.n_total = n
            ens.gen_beta = [float(v) for v in d["gen_beta"]]
            ens.gen_logz = [float(v) for v in d["gen_logz"]]
            ens.gen_n_local = [int(v) for v in d["gen_n_local"]]
            ens.gen_n = [int(v) * ens.world for v in d["gen_n_local"]]
            ens._sync_gens()
            ens.rebuild_mixture()                           # the cached log-mixture column is derived data
        self.ensemble = ens
        st = self.state
        for key in list(st._history):
            if key != "blobs" and ("hist_" + key) in d:
                vals = d["hist_" + key]
                st._history[key] = [int(v) if key in ("iter", "calls", "steps") else float(v) for v in vals]
        defaults = {"iter": 0, "calls": 0, "beta": 0.0, "logz": 0.0, "steps": 0, "acceptance": 0.0,
                    "efficiency": 0.0, "ess": None, "cv": None}                          # core.py:296-309
        for key, default in defaults.items():
            v = float(d["cur_" + key]) if ("cur_" + key) in d else float("nan")
            if np.isnan(v):
                st.set_current(key, default)
            else:
                st.set_current(key, int(v) if key in ("iter", "calls", "steps") else v)
        for key in ("u", "logl"):
            if ("cur_" + key) in d:
                st.set_current(key, torch.as_tensor(d["cur_" + key]).to(self.device))
        if "cur_assignments" in d:
            st.set_current("assignments", d["cur_assignments"])
        if self.clusterer is not None:
            self.clusterer.load_state_dict(d, self.device)
        self.assign = (torch.as_tensor(d["core_assign"]).to(self.device) if "core_assign" in d else None)
        self.run_epoch = int(d["run_epoch"]) if "run_epoch" in d else 0
        st.set_current("x", None)
        self.n_total = int(d["n_total"])
        self.rng.seed = int(d["rng_seed"])                  # continue the same counter-based stream
        self.rng.set_epoch(self.run_epoch)
        self._weights = None

    # -- results ----------------------------------------------------------------------------------
    def compute_posterior(self, resample=False, return_blobs=False, trim_importance_weights=True,
                          return_logw=False, ess_trim=0.99, bins_trim=1000):
        ens = self.ensemble
        n = ens.n_total
        k = self.k
        stats = k.probe(ens, 1.0, torch.zeros(16, dtype=F64, device=self.device))
        w = torch.empty(n, dtype=F64, device=self.device)
        k.weights(ens, 1.0, stats, w)                     # core.py:197-199
        logw = None
        if return_logw:
            lw = torch.empty(n, dtype=F64, device=self.device)
            k.weights(ens, 1.0, stats, lw, log=True)
            logw = lw.cpu().numpy()
        u, logl = ens.u[:n], ens.logl[:n]
        idx = None
        if trim_importance_weights:                        # core.py:210-220
            idx, w = k.trim(w, n, ess=ess_trim, bins=bins_trim, n_global=ens.n_total_global)
            u, logl = u[idx], logl[idx]
        if resample and self.comm.on:                      # core.py:222-231 over the global (trimmed) weight vector
            m_loc = int(w.numel())
            m = k.g_int(m_loc)
            bounds = self.generation_bounds()
            seg_begin = torch.searchsorted(idx, bounds) if idx is not None else bounds
            h = k.cdf_x(w, m_loc, seg_begin, m, "post_cdf")
            pos = torch.empty(m, dtype=torch.int64, device=self.device)
            k.search_x(h, None, m, pos, systematic=True, u0=self.rng.resample_u0())
            own = torch.nonzero(pos >= 0).flatten()        # positions whose ancestor this rank stores
            src = pos[own]
            x = self.transform_to_x(u[src].contiguous())
            allx, alll, allp = (self.comm.allgather_rows(t.contiguous()) for t in (x, logl[src], own))
            order = torch.argsort(allp)                    # every rank returns the sample in position order
            x, logl = allx[order], alll[order]
            w = torch.full((m,), 1.0 / m, dtype=F64, device=self.device)
            k.check_pending()
            out = tuple(self._to_host(t) for t in (x, w, logl))
            if logw is not None:
                logw = self.comm.allgather_rows(torch.as_tensor(logw).to(self.device)).cpu().numpy()
            return out + (logw,) if return_logw else out
        if resample:                                       # core.py:222-231
            m = int(w.numel())
            cdf = k.cdf(w, m, "post_cdf")
            idx = torch.empty(m, dtype=torch.int64, device=self.device)
            k.systematic(cdf, m, self.rng.resample_u0(), m, idx)
            u, logl = u[idx], logl[idx]
            w = torch.full((m,), 1.0 / m, dtype=F64, device=self.device)
        x = self.transform_to_x(u.contiguous())
        if self.comm.on:                                   # every rank returns the global (rank-major) sample
            x, w, logl = (self.comm.allgather_rows(t.contiguous()) for t in (x, w, logl))
            if logw is not None:
                logw = self.comm.allgather_rows(torch.as_tensor(logw).to(self.device)).cpu().numpy()
        out = tuple(self._to_host(t) for t in (x, w, logl))
        return out + (logw,) if return_logw else out

    @staticmethod
    def _to_host(t: torch.Tensor) -> np.ndarray:
        """Device -> host through page-locked memory (the weighted posterior of a 2^20-particle run is
        ~600 MB; a pageable ``.cpu()`` moves it at ~2 GB/s).  The numpy array owns the pinned block; torch's
        host allocator recycles it once the array is released."""
        if t.numel() < (1 << 16):
            return t.cpu().numpy()
        host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        host.copy_(t, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host.numpy()

    def compute_evidence(self):
        return self.state.raw("logz"), self.logz_err
