"""Sampler configuration: same fields, defaults, validation messages and algorithm constants as
the reference (tempest/config.py:10-242), plus the restrictions of the CUDA path."""
from __future__ import annotations

import warnings
from dataclasses import dataclass, fields
from pathlib import Path
from typing import Any, Callable, List, Optional, Union

# Algorithm constants -- tempest/config.py:233-242 (part of results parity)
BETA_TOLERANCE: float = 1e-4
BETA_RTOL: float = 1e-8
ESS_TOLERANCE: float = 0.01
METRIC_ATOL: float = 0.5
METRIC_ATOL_CV: float = 0.01
DOF_FALLBACK: float = 1e6
TRIM_ESS: float = 0.99
TRIM_BINS: int = 1000
MAX_BISECTION_ITERATIONS: int = 200  # tempest/steps/reweight.py:121


@dataclass(frozen=True)
class SamplerConfig:
    """Immutable, validated configuration (reference: tempest/config.py:10-185)."""

    prior_transform: Callable
    log_likelihood: Callable
    n_dim: int
    n_particles: Optional[int] = None
    ess_ratio: float = 2.0
    volume_variation: Optional[float] = None
    log_likelihood_args: Optional[list] = None
    log_likelihood_kwargs: Optional[dict] = None
    vectorize: bool = False
    blobs_dtype: Optional[str] = None
    periodic: Optional[List[int]] = None
    reflective: Optional[List[int]] = None
    pool: Optional[Union[int, Any]] = None
    clustering: bool = True
    normalize: bool = True
    cluster_every: int = 1
    split_threshold: float = 1.0
    n_max_clusters: Optional[int] = None
    sample: str = "tpcn"
    n_steps: Optional[int] = None
    n_max_steps: Optional[int] = None
    resample: str = "mult"
    output_dir: Optional[Path] = None
    output_label: Optional[str] = None
    random_state: Optional[int] = None

    def __post_init__(self) -> None:
        put = lambda k, v: object.__setattr__(self, k, v)  # noqa: E731 (frozen dataclass)
        if not isinstance(self.n_dim, int):
            raise ValueError(f"n_dim must be int, got {type(self.n_dim).__name__}")
        if self.output_dir is None:
            put("output_dir", Path("states"))
        elif isinstance(self.output_dir, str):
            put("output_dir", Path(self.output_dir))
        if self.output_label is None:
            put("output_label", "ps")
        if self.n_particles is None:
            put("n_particles", 2 * self.n_dim)  # config.py:75-76
        if self.n_steps is None or self.n_steps <= 0:
            put("n_steps", 1)  # config.py:80-81
        if self.n_max_steps is None or self.n_max_steps <= 0:
            put("n_max_steps", 20 * self.n_steps)  # config.py:83-84
        self.validate()
        if self.volume_variation is not None and self.n_particles < self.n_dim + 1:
            warnings.warn(
                f"For dynamic mode, n_particles ({self.n_particles}) "
                f"should be >= n_dim + 1 ({self.n_dim + 1}) for reliable results. "
                f"Volume variation calculation may be inaccurate.",
                UserWarning,
                stacklevel=2,
            )

    def validate(self) -> None:
        """Collect every problem and raise one ValueError (format of config.py:181-185)."""
        bad: List[str] = []
        if not callable(self.prior_transform):
            bad.append("prior_transform must be callable")
        if not callable(self.log_likelihood):
            bad.append("log_likelihood must be callable")
        if not isinstance(self.n_dim, int) or self.n_dim <= 0:
            bad.append(f"n_dim must be positive int, got {self.n_dim}")
        if not isinstance(self.n_particles, int):
            bad.append(f"n_particles must be int, got {type(self.n_particles)}")
        if self.n_particles <= 0:
            bad.append(f"n_particles must be positive integer, got {self.n_particles}")
        if not isinstance(self.ess_ratio, (int, float)):
            bad.append(f"ess_ratio must be numeric, got {type(self.ess_ratio)}")
        if self.ess_ratio <= 0:
            bad.append(f"ess_ratio must be positive, got {self.ess_ratio}")
        if self.volume_variation is not None:
            if not isinstance(self.volume_variation, (int, float)):
                bad.append(f"volume_variation must be numeric or None, got {type(self.volume_variation)}")
            elif self.volume_variation <= 0:
                bad.append(f"volume_variation ({self.volume_variation}) must be positive")
        if self.sample not in ["tpcn", "rwm"]:
            bad.append(f"Invalid sampler '{self.sample}': must be 'tpcn' or 'rwm'")
        if self.resample not in ["mult", "syst"]:
            bad.append(f"Invalid resample '{self.resample}': must be 'mult' or 'syst'")
        if self.vectorize and self.blobs_dtype is not None:
            bad.append("Cannot vectorize likelihood with blobs")
        if self.periodic is not None and self.reflective is not None:
            both = set(self.periodic).intersection(set(self.reflective))
            if both:
                bad.append(f"Parameters cannot be both periodic and reflective: {both}")
        for label, lst in (("periodic", self.periodic), ("reflective", self.reflective)):
            if lst is not None and not all(isinstance(i, int) and 0 <= i < self.n_dim for i in lst):
                bad.append(f"{label} indices must be integers in [0, {self.n_dim - 1}], got {lst}")
        if not isinstance(self.output_dir, Path):
            bad.append(f"output_dir must be Path, got {type(self.output_dir)}")
        if self.output_label is not None and not isinstance(self.output_label, str):
            bad.append(f"output_label must be str or None, got {type(self.output_label)}")
        if bad:
            raise ValueError("Configuration validation failed:\n" + "\n".join(f"  - {m}" for m in bad))

    def get_target_metric(self) -> float:
        if self.volume_variation is not None:
            return self.volume_variation
        return self.ess_ratio * self.n_particles

    def to_dict(self) -> dict:
        out = {f.name: getattr(self, f.name) for f in fields(self)}
        out["output_dir"] = str(self.output_dir)
        return out
