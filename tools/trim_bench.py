"""Times the sub-steps of Kernels.trim on realistic late-stage weights (C4 ensemble at full size)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import tempest_b200 as tp
from tempest_b200 import _lib
s = tp.Sampler(tp.UniformPrior(-10.0, 10.0, 10), tp.Rosenbrock(10), 10, n_particles=1 << 20, vectorize=True,
               clustering=False, random_state=20261018)
core = s._core; core._initialize_fresh(); core.n_total = 4096
for _ in range(30):
    core.execute_iteration(export=False)
k, ens = core.k, core.ensemble
n = ens.n_total
k.probe(ens, 0.3)
w0 = torch.empty(n, dtype=torch.float64, device=core.device); k.weights(ens, 0.3, k.probe_out, w0)
# monkeypatch timing around the library calls
times = {}
lib = k.lib
def wrap(name):
    fn = getattr(lib, name)
    def f(*a):
        torch.cuda.synchronize(); t = time.perf_counter(); r = fn(*a); torch.cuda.synchronize()
        times[name] = times.get(name, 0.0) + (time.perf_counter() - t) * 1e3; times[name + "#"] = times.get(name + "#", 0) + 1
        return r
    setattr(lib, name, f)
for nm in ("tb_normalize_inplace", "tb_binade_hist", "tb_compact_ge", "tb_select_pair", "tb_masked_sums"):
    wrap(nm)
for rep in range(3):
    w = w0.clone(); times.clear()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    idx, wt = k.trim(w, n)
    torch.cuda.synchronize(); tot = (time.perf_counter() - t0) * 1e3
    print(f"trim total {tot:.2f} ms (with per-call syncs) n={n} n_trim={idx.numel()} exact={k.last_trim['n_exact']}", {a: round(b, 3) for a, b in times.items()})
