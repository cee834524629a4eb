"""Micro-benchmark of the ESS probe / weights / cdf / moments kernels on a synthetic persistent
ensemble: prints achieved GB/s against the algorithmic bytes of each kernel."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from tempest_b200 import _lib
from tempest_b200.ensemble import PersistentEnsemble, ptr, stream_ptr
from tempest_b200.steps import Kernels

n_gen = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
T = int(sys.argv[2]) if len(sys.argv) > 2 else 32
d = 10
dev = torch.device("cuda:0")
k = Kernels(dev)
ens = PersistentEnsemble(d, dev)
g = torch.Generator(device=dev).manual_seed(20261018)
betas = [0.0] * 3 + list(np.geomspace(1e-4, 1.0, T - 3))
for t in range(T):
    u = torch.rand((n_gen, d), dtype=torch.float64, device=dev, generator=g)
    chi = (torch.randn((n_gen, d), dtype=torch.float64, device=dev, generator=g) ** 2).sum(1)
    ens.append(u, -0.5 * chi * 4.0, betas[t], -0.3 * t)
n = ens.n_total
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
res = {}
ms = timeit(lambda: k.probe(ens, 0.37)); res["probe"] = (ms, 16 * n / ms / 1e6)
w = torch.empty(n, dtype=torch.float64, device=dev)
k.probe(ens, 0.37)
ms = timeit(lambda: k.weights(ens, 0.37, k.probe_out, w)); res["weights"] = (ms, 24 * n / ms / 1e6)
ms = timeit(lambda: k.next_beta(ens, 0.2, 2.0 * n_gen, 0), reps=5); h = k.ws.f64("nb_res", 16).cpu().numpy()
res["next_beta"] = (ms, 16 * n * h[6] / ms / 1e6, int(h[6]))
ms = timeit(lambda: k.cdf(w, n)); res["cdf_exact"] = (ms, 16 * n / ms / 1e6)
ms = timeit(lambda: k.volume_variation(ens.u, w, n, d), reps=5); res["volume_variation(3 passes)"] = (ms, (3 * 8 * d + 3 * 8) * n / ms / 1e6)
dr = torch.rand(n_gen, dtype=torch.float64, device=dev, generator=g)
idx = torch.empty(n_gen, dtype=torch.int64, device=dev)
cdf = k.cdf(w, n)
ms = timeit(lambda: k.search_right(cdf, n, dr, idx)); res["search_right"] = (ms, None)
for name, v in res.items():
    print(f"{name:28s} {v[0]:8.3f} ms  {('%.0f GB/s' % v[1]) if v[1] else ''} {v[2:] if len(v) > 2 else ''}")
print(json.dumps({"n": n, "T": T, **{k_: v for k_, v in res.items()}}))
