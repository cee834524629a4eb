"""configs[2]: 50-D AR(1) Gaussian, N = 2^18 (or argv[1]), pCN; analytic logZ = -50 log 20."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import tempest_b200 as tp
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
d = 50
s = tp.Sampler(tp.UniformPrior(-10.0, 10.0, d), tp.GaussianLikelihood.ar1(d, 0.5), d, n_particles=n, vectorize=True,
               clustering=False, random_state=3)
core = s._core; core.profile = True
torch.cuda.synchronize(); t0 = time.perf_counter()
s.run(progress=False)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
steps = s.state.get_history("steps")
print(json.dumps(dict(n=n, d=d, T=len(steps), seconds=round(dt, 2), logz=s.evidence()[0], exact=-d * np.log(20.0),
                      steps_total=int(steps.sum()), ms_per_step=round(core.stage_ms.get("mutate", 0) / max(1, steps[3:].sum()), 3),
                      stages={k: round(v, 1) for k, v in core.stage_ms.items()})))
