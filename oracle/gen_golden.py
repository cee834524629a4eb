"""Generate golden fixtures from the UNMODIFIED reference (run in the build container only).

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden.py [--ref /root/reference]

For every case below the reference ``tempest.Sampler`` is constructed with registry callables
(plain numpy functions, ``tempest_b200/registry.py``), the legacy global MT19937 stream is
seeded with ``np.random.seed(seed)`` -- the reference never seeds a fresh run itself
(core.py:314-315) -- and ``run()`` is called.  Nothing is patched.  The fixture keeps the
whole persistent ensemble (u, x, logl per generation), the per-generation scalars and the
``posterior()`` / ``evidence()`` outputs.  ``tests/test_oracle_golden.py`` replays
``oracle/ps_oracle.py`` on ``LegacyStream(seed)`` against these files; the GPU tests then
compare the CUDA path with the oracle on the tapes the oracle records.

Test infrastructure only; never imported by ``tempest_b200``.
"""

from __future__ import annotations

import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tempest_b200.registry import (  # noqa: E402
    GaussianLikelihood, IsotropicMixture, Rosenbrock, TwinShells, UniformPrior)


def cases():
    """name -> (prior, likelihood, sampler kwargs, n_total, seed)."""
    ar1 = GaussianLikelihood.ar1(4, 0.5)
    return {
        # README likelihood, small ensemble, defaults otherwise (tpCN + multinomial)
        "rosen10_n64_tpcn_mult": (
            UniformPrior(-10.0, 10.0, 10), Rosenbrock(10),
            dict(n_dim=10, n_particles=64, clustering=False), 512, 11),
        # random-walk Metropolis + systematic resampling + periodic/reflective boundaries
        "gauss4_n32_rwm_syst_bc": (
            UniformPrior(-6.0, 6.0, 4), ar1,
            dict(n_dim=4, n_particles=32, clustering=False, sample="rwm", resample="syst",
                 periodic=[0], reflective=[1]), 256, 12),
        # non power-of-two ensemble (exercises the warm-up ESS == target rounding)
        "mix2_n21_tpcn_mult": (
            UniformPrior(-10.0, 10.0, 2), IsotropicMixture.four_corners(2),
            dict(n_dim=2, n_particles=21, clustering=False, n_steps=2), 128, 13),
        # dynamic (volume-variation) reweighting mode, reweight.py:427-495
        "shell4_n48_dynamic": (
            UniformPrior(-6.0, 6.0, 4), TwinShells(4),
            dict(n_dim=4, n_particles=48, clustering=False, volume_variation=0.5), 192, 14),
    }


def run_reference(ref_root, prior, like, kwargs, n_total, seed):
    sys.path.insert(0, ref_root)
    import tempest as tp  # the reference

    np.random.seed(seed)
    s = tp.Sampler(prior_transform=prior, log_likelihood=like, vectorize=True, **kwargs)
    s.run(n_total=n_total, progress=False)
    st = s.state
    out = {}
    for key in ("u", "x", "logl"):
        out[key] = np.asarray(st.get_history(key))
    for key in ("iter", "logz", "calls", "steps", "efficiency", "ess", "cv", "acceptance", "beta"):
        out["h_" + key] = np.asarray(st.get_history(key), dtype=float)
    logz, _ = s.evidence()
    out["final_logz"] = np.array(logz)
    x, w, l, logw = s.posterior(return_logw=True)
    out.update(post_x=x, post_w=w, post_logl=l, post_logw=logw)
    x2, w2, l2 = s.posterior(trim_importance_weights=False)
    out.update(post_w_untrimmed=w2)
    out["seed"] = np.array(seed)
    out["n_total"] = np.array(n_total)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    for name, (prior, like, kw, n_total, seed) in cases().items():
        if a.only and a.only != name:
            continue
        out = run_reference(a.ref, prior, like, kw, n_total, seed)
        path = os.path.join(a.out, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: T={len(out['h_beta'])} logz={float(out['final_logz']):.6f} "
              f"steps={out['h_steps'].sum():.0f} -> {path} ({os.path.getsize(path)/1024:.0f} KiB)")


if __name__ == "__main__":
    main()
