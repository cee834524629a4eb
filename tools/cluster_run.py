"""Config C2: 2-D four-mode mixture, clustering=True, analytic logZ = -log 400.
usage: python tools/cluster_run.py [n_particles] [runs]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import tempest_b200 as tp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 2
d = 2
for r in range(runs):
    s = tp.Sampler(tp.UniformPrior(-10.0, 10.0, d), tp.IsotropicMixture.four_corners(d), d, n_particles=n,
                   vectorize=True, clustering=True, random_state=20261018 + r)
    core = s._core
    core.profile = True
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s.run(n_total=4096, progress=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    T = s.state.get_history_length()
    logz = s.evidence()[0]
    stages = getattr(core, "stage_ms", {})
    print(json.dumps(dict(run=r, n=n, T=T, seconds=round(dt, 3), it_per_s=round(T / dt, 2), logz=logz,
                          logz_exact=-np.log(400.0), K=core.clusterer.n_clusters_, fits=core.clusterer.n_fits,
                          calls=int(s.state.raw("calls")), steps=[int(v) for v in s.state.get_history("steps")],
                          stages_ms={k: round(v, 1) for k, v in sorted(stages.items(), key=lambda kv: -kv[1])})))
