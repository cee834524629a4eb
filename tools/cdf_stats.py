"""Structure of the weight vectors the exact cumulative sum sees in a real C4 run (design input for tb_resample.cu):
binade crossings of the running sum, how violent they are (crossing element / running sum), zeros, ties."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import tempest_b200 as tp
from tempest_b200.steps import Kernels
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
d = 10
s = tp.Sampler(tp.UniformPrior(-10.0, 10.0, d), tp.Rosenbrock(d), d, n_particles=n, vectorize=True, clustering=False,
               random_state=20261018)
core = s._core
core._initialize_fresh(); core.n_total = 4096
it = 0
while core._not_termination():
    core.execute_iteration(export=False)
    it += 1
    if it in (6, 12, 20, 28, 34, 36):
        ens = core.ensemble
        k = core.k
        k.probe(ens, float(core.state.raw("beta")))
        w = k.weights(ens, float(core.state.raw("beta")), k.probe_out, core.weights_buffer()).clone()
        w /= w.sum()
        p = w.cpu().numpy()
        c = np.cumsum(p)
        nz = c > 0
        first = int(np.argmax(nz)) if nz.any() else len(c)
        e = np.frexp(c[first:])[1]
        cross = np.nonzero(np.diff(e) != 0)[0] + 1 + first          # index of the element that crossed
        prev = c[cross - 1]
        ratio = p[cross] / prev
        jump = np.diff(e)[cross - 1 - first]
        tiles = np.unique(cross // 1024)
        print(f"it {it} beta {core.state.raw('beta'):.4g} N_total {len(p)}: zeros {np.count_nonzero(p == 0)} first nonzero {first} "
              f"crossings {len(cross)} in {len(tiles)} tiles; multi-binade jumps {np.count_nonzero(jump > 1)}; "
              f"violent (elem >= running sum) {np.count_nonzero(ratio >= 1.0)}; elem > 1e-3 sum {np.count_nonzero(ratio > 1e-3)}; "
              f"min exp {e.min()} ; crossings below 2^-960: {np.count_nonzero(np.frexp(prev)[1] < -960)}", flush=True)
        # per-tile: tiles with > 1 crossing
        cnt = np.bincount(cross // 1024)
        print("   tiles with >1 crossing:", np.count_nonzero(cnt > 1), " max crossings in a tile:", cnt.max(),
              " crossings in first 5% of positions:", np.count_nonzero(cross < 0.05 * len(p)),
              " subnormal elements:", np.count_nonzero((p > 0) & (p < 2.3e-308)), flush=True)
        del p, c, w
