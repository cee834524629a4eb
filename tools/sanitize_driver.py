"""Small end-to-end workload for compute-sanitizer (tools/sanitize.sh): every stage of a PS iteration at small
sizes -- cooperative ESS search with its grid-wide fold, trim, exact cdf (incl. hard tiles), searches, gather,
moments, persistent Metropolis kernel (fast and wide bodies), clustering kernels, posterior()."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import tempest_b200 as tp

def run(prior, like, d, n, iters, **kw):
    s = tp.Sampler(prior, like, d, n_particles=n, vectorize=True, random_state=5, **kw)
    for _ in range(iters):
        s.sample()
    x, w, l = s.posterior()
    assert np.isfinite(l).all() and abs(w.sum() - 1.0) < 1e-9
    return s

run(tp.UniformPrior(-10.0, 10.0, 10), tp.Rosenbrock(10), 10, 2048, 7, clustering=False)
run(tp.UniformPrior(-6.0, 6.0, 4), tp.GaussianLikelihood.ar1(4, 0.5), 4, 512, 6, clustering=False, sample="rwm",
    resample="syst", periodic=[0], reflective=[1])
run(tp.UniformPrior(-10.0, 10.0, 2), tp.IsotropicMixture.four_corners(2), 2, 256, 6, clustering=True)
run(tp.UniformPrior(-6.0, 6.0, 24), tp.TwinShells(24), 24, 256, 5, clustering=False)
print("sanitize driver ok")
