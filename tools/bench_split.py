import os, sys, time, gc
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import tempest_b200 as tp
d = 10
def new():
    return tp.Sampler(tp.UniformPrior(-10.0, 10.0, d), tp.Rosenbrock(d), d, n_particles=1 << 20, vectorize=True,
                      clustering=False, random_state=20261018)
w = new(); w.run(progress=False); del w
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s = new(); core = s._core; core._initialize_fresh(); core.n_total = 4096
    torch.cuda.synchronize(); t1 = time.perf_counter()
    marks = []
    while core._not_termination():
        core.execute_iteration(export=False)
        if core.state.raw("iter") in (1, 3, 10, 20, 30):
            torch.cuda.synchronize(); marks.append((core.state.raw("iter"), round(time.perf_counter() - t1, 4)))
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(dict(create=round(t1 - t0, 4), run=round(t2 - t1, 4), marks=marks, T=core.state.get_history_length()))
    if rep == 1:
        gc.collect()
