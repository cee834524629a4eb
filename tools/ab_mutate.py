"""A/B timing of the mutation stage: one C4 run (N walkers, 10-D Rosenbrock) per library variant given in
TEMPEST_B200_LIB; prints total mutate ms, MCMC steps and ms/step."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import tempest_b200 as tp
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
d = 10
out = []
for rep in range(2):
    s = tp.Sampler(tp.UniformPrior(-10.0, 10.0, d), tp.Rosenbrock(d), d, n_particles=n, vectorize=True,
                   clustering=False, random_state=20261018)
    core = s._core
    core.profile = True
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s.run(progress=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    steps = int(sum(s.state.get_history("steps")[3:]))
    out.append(dict(lib=os.path.basename(os.environ.get("TEMPEST_B200_LIB", "default")), rep=rep, seconds=round(dt, 3),
                    T=s.state.get_history_length(), steps=steps, mutate_ms=round(core.stage_ms.get("mutate", 0.0), 1),
                    ms_per_step=round(core.stage_ms.get("mutate", 0.0) / steps, 4), logz=round(s.evidence()[0], 4)))
print(json.dumps(out[-1]))
