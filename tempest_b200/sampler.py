"""Public facade: the reference's ``tempest.Sampler`` surface (tempest/sampler.py:12-406) over
the device Persistent Sampling loop.  Signature, defaults, return shapes, property names and
``ValueError`` messages follow the reference; everything underneath runs in libtempest_b200.
"""
from __future__ import annotations

from pathlib import Path
from typing import Optional, Union

from .config import SamplerConfig
from .core import SamplerCore


class FunctionWrapper:
    """Bind extra args/kwargs to the likelihood (tempest/tools.py:270-309)."""

    def __init__(self, f, args, kwargs):
        self.f = f
        self.args = [] if args is None else args
        self.kwargs = {} if kwargs is None else kwargs

    def __call__(self, x):
        return self.f(x, *self.args, **self.kwargs)


class Sampler:
    """Drop-in for ``tempest.Sampler`` on the vectorised hot path (sampler.py:22-49)."""

    def __init__(
        self,
        prior_transform: callable,
        log_likelihood: callable,
        n_dim: int,
        n_particles: Optional[int] = None,
        ess_ratio: float = 2.0,
        volume_variation: Optional[float] = None,
        log_likelihood_args: Optional[list] = None,
        log_likelihood_kwargs: Optional[dict] = None,
        vectorize: bool = False,
        blobs_dtype: Optional[str] = None,
        periodic: Optional[list] = None,
        reflective: Optional[list] = None,
        pool: Optional[Union[int, object]] = None,
        clustering: bool = True,
        normalize: bool = True,
        cluster_every: int = 1,
        split_threshold: float = 1.0,
        n_max_clusters: Optional[int] = None,
        sample: str = "tpcn",
        n_steps: Optional[int] = None,
        n_max_steps: Optional[int] = None,
        resample: str = "mult",
        output_dir: Optional[str] = None,
        output_label: Optional[str] = None,
        random_state: Optional[int] = None,
    ):
        wrapped = FunctionWrapper(log_likelihood, log_likelihood_args, log_likelihood_kwargs)
        config = SamplerConfig(
            prior_transform=prior_transform, log_likelihood=wrapped, n_dim=n_dim, n_particles=n_particles,
            ess_ratio=ess_ratio, volume_variation=volume_variation, log_likelihood_args=log_likelihood_args,
            log_likelihood_kwargs=log_likelihood_kwargs, vectorize=vectorize, blobs_dtype=blobs_dtype,
            periodic=periodic, reflective=reflective, pool=pool, clustering=clustering, normalize=normalize,
            cluster_every=cluster_every, split_threshold=split_threshold, n_max_clusters=n_max_clusters,
            sample=sample, n_steps=n_steps, n_max_steps=n_max_steps, resample=resample, output_dir=output_dir,
            output_label=output_label, random_state=random_state,
        )
        self._core = SamplerCore(config)
        self.state = self._core.state          # tests reach for sampler.state (sampler.py:160-161)

    # -- running ----------------------------------------------------------------------------
    def run(self, n_total: int = 4096, progress: bool = True,
            resume_state_path: Union[str, Path, None] = None, save_every: Optional[int] = None):
        return self._core.run_sampling(n_total=n_total, progress=progress,
                                       resume_state_path=resume_state_path, save_every=save_every)

    def sample(self, save_every: Optional[int] = None, t0: int = 0) -> dict:
        if self.state.raw("iter") is None:
            self._core._initialize_fresh()
        return self._core.execute_iteration(save_every=save_every, t0=t0)

    def posterior(self, resample: bool = False, return_blobs: bool = False,
                  trim_importance_weights: bool = True, return_logw: bool = False,
                  ess_trim: float = 0.99, bins_trim: int = 1000) -> tuple:
        return self._core.compute_posterior(
            resample=resample, return_blobs=return_blobs, trim_importance_weights=trim_importance_weights,
            return_logw=return_logw, ess_trim=ess_trim, bins_trim=bins_trim)

    def evidence(self) -> tuple:
        return self._core.compute_evidence()

    def results(self):
        return self.state.compute_results()

    def save_state(self, path):                          # sampler.py:278-287
        self._core.save_sampler_state(Path(path))

    def load_state(self, path):                          # sampler.py:289-298
        self._core.load_sampler_state(Path(path))

    # -- read-only properties (sampler.py:313-406) ---------------------------------------------
    n_dim = property(lambda self: self._core.config.n_dim)
    n_particles = property(lambda self: self._core.config.n_particles)
    ess_ratio = property(lambda self: self._core.config.ess_ratio)
    volume_variation = property(lambda self: self._core.config.volume_variation)
    n_steps = property(lambda self: self._core.config.n_steps)
    n_max_steps = property(lambda self: self._core.config.n_max_steps)
    n_total = property(lambda self: self._core.n_total or None)
    resample = property(lambda self: self._core.config.resample)
    clustering = property(lambda self: self._core.config.clustering)
    vectorize = property(lambda self: self._core.config.vectorize)
    output_dir = property(lambda self: self._core.config.output_dir)
    output_label = property(lambda self: self._core.config.output_label)
    random_state = property(lambda self: self._core.config.random_state)
    periodic = property(lambda self: self._core.config.periodic)
    reflective = property(lambda self: self._core.config.reflective)
    beta = property(lambda self: self.state.get_current("beta"))
    logz = property(lambda self: self.state.get_current("logz"))
    ess = property(lambda self: self.state.get_current("ess"))
    cv = property(lambda self: self.state.get_current("cv"))
