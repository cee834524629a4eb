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
        # hierarchical-GMM clustering on (defaults): four separated modes in 2-D (config C2, small)
        "mix2_n64_clustered": (
            UniformPrior(-10.0, 10.0, 2), IsotropicMixture.four_corners(2),
            dict(n_dim=2, n_particles=64, clustering=True), 256, 21),
        # clustering with a cluster cap (core.py:59-69), a lower split threshold, refit every 2nd iteration
        "mix3_n96_clustered_cap": (
            UniformPrior(-10.0, 10.0, 3), IsotropicMixture.four_corners(3),
            dict(n_dim=3, n_particles=96, clustering=True, cluster_every=2, n_max_clusters=3,
                 split_threshold=0.5), 256, 22),
    }


def cluster_cases():
    """name -> (X[n,d], w[n], HierarchicalGaussianMixture kwargs): stand-alone fits (cluster.py)."""
    rs = np.random.RandomState(77)

    def blobs(n, d, k, spread, lo=0.2, hi=0.8):
        c = rs.rand(k, d) * (hi - lo) + lo
        lab = rs.randint(0, k, n)
        return c[lab] + spread * rs.randn(n, d)

    out = {}
    out["cluster_blobs2d"] = (blobs(600, 2, 4, 0.02), rs.rand(600) ** 2, dict(normalize=True))
    out["cluster_blobs5d"] = (blobs(500, 5, 3, 0.03), rs.rand(500) ** 3, dict(normalize=True))
    out["cluster_single10d"] = (blobs(300, 10, 1, 0.05), rs.rand(300), dict(normalize=True))
    out["cluster_raw3d_cap"] = (blobs(400, 3, 3, 0.01), rs.rand(400),
                                dict(normalize=False, max_iterations=1, min_points=12, threshold_modifier=0.5))
    x = blobs(256, 2, 2, 0.04)
    x[:, 1] = 0.5 + 1e-3 * (x[:, 1] - 0.5)      # nearly degenerate second axis
    out["cluster_thin2d"] = (x, np.ones(256), dict(normalize=True))
    return out


def run_reference_cluster(ref_root, x, w, kw):
    sys.path.insert(0, ref_root)
    from tempest.cluster import GaussianMixture, HierarchicalGaussianMixture  # the reference

    h = HierarchicalGaussianMixture(**kw).fit(x, w)
    rs = np.random.RandomState(3)
    y = rs.rand(512, x.shape[1])
    out = dict(x=x, w=w, y=y, labels=h.labels_, centres=np.array(h.cluster_centers_),
               covs=np.array(h.cluster_covariances_), weights=h.cluster_weights_,
               n_clusters=np.array(h.n_clusters_), predict_y=h.predict(y), predict_x=h.predict(x))
    for k in (1, 2):
        g = GaussianMixture(n_components=k, random_state=42).fit(x, w)
        out.update({f"gmm{k}_weights": g.weights_, f"gmm{k}_means": g.means_, f"gmm{k}_covs": g.covariances_,
                    f"gmm{k}_n_iter": np.array(g.n_iter_), f"gmm{k}_bic": np.array(g.bic(x)),
                    f"gmm{k}_labels": g.predict(x), f"gmm{k}_lower": np.array(g.lower_bound_)})
    return out


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
    main_cluster(a.ref, a.out, a.only)


def main_cluster(ref, out_dir, only=None):
    for name, (x, w, kw) in cluster_cases().items():
        if only and only != name:
            continue
        out = run_reference_cluster(ref, x, w, kw)
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: K={int(out['n_clusters'])} gmm2 iters={int(out['gmm2_n_iter'])} -> {path} "
              f"({os.path.getsize(path)/1024:.0f} KiB)")


if __name__ == "__main__":
    main()
