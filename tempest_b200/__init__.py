"""tempest_b200 -- B200-native Persistent Sampling inner loop behind the ``tempest.Sampler`` API.

    import tempest_b200 as tp
    sampler = tp.Sampler(tp.UniformPrior(-10, 10, 10), tp.Rosenbrock(10), n_dim=10,
                         vectorize=True, clustering=False)
    sampler.run(); x, w, logl = sampler.posterior(); logz, _ = sampler.evidence()

Importing the package never touches CUDA; constructing a ``Sampler`` loads
``lib/libtempest_b200.so`` and fails loudly if it (or a CUDA device) is missing.
"""
from .registry import (GaussianLikelihood, IsotropicMixture, Rosenbrock, TwinShells,  # noqa: F401
                       UniformPrior)

__version__ = "0.1.0"


def device_callable(fn):
    """Mark a prior_transform / log_likelihood as operating on CUDA fp64 torch tensors (callables.py)."""
    fn._tb_device = True
    return fn


def __getattr__(name):
    if name == "Sampler":
        from .sampler import Sampler

        return Sampler
    raise AttributeError(name)
