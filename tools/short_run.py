"""Short run for ncu captures: N particles, stops after `iters` PS iterations."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import tempest_b200 as tp
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 8
d = 10
s = tp.Sampler(tp.UniformPrior(-10.0, 10.0, d), tp.Rosenbrock(d), d, n_particles=n, vectorize=True,
               clustering=False, random_state=20261018)
core = s._core
core._initialize_fresh()
core.n_total = 4096
for _ in range(iters):
    core.execute_iteration(export=False)
torch.cuda.synchronize()
print("done", core.state.raw("beta"), core.state.raw("steps"))
