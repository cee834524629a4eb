"""Where does posterior() spend its time after a full C4 run?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, numpy as np
import tempest_b200 as tp
d = 10
for rep in range(2):
    s = tp.Sampler(tp.UniformPrior(-10.0, 10.0, d), tp.Rosenbrock(d), d, n_particles=1 << 20, vectorize=True,
                   clustering=False, random_state=20261018)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s.run(progress=False)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    core = s._core
    ens = core.ensemble; k = core.k; n = ens.n_total
    stats = k.probe(ens, 1.0, torch.zeros(16, dtype=torch.float64, device=core.device))
    w = torch.empty(n, dtype=torch.float64, device=core.device)
    k.weights(ens, 1.0, stats, w)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    idx, wt = k.trim(w, n)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    u = ens.u[:n][idx]; logl = ens.logl[:n][idx]
    torch.cuda.synchronize(); t4 = time.perf_counter()
    x = core.transform_to_x(u.contiguous())
    torch.cuda.synchronize(); t5 = time.perf_counter()
    xh = x.cpu().numpy(); wh = wt.cpu().numpy(); lh = logl.cpu().numpy()
    t6 = time.perf_counter()
    x2, w2, l2 = s.posterior()
    t7 = time.perf_counter()
    print(dict(run=round(t1 - t0, 3), weights=round(t2 - t1, 4), trim=round(t3 - t2, 4), gather=round(t4 - t3, 4),
               transform=round(t5 - t4, 4), d2h=round(t6 - t5, 4), posterior_call=round(t7 - t6, 4), n_post=len(wh)))
