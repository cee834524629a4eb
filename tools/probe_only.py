"""Synthetic persistent ensemble + a few ESS probes / one next-beta search: ncu target."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from tempest_b200.ensemble import PersistentEnsemble
from tempest_b200.steps import Kernels
n_gen, T, d = (int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20), (int(sys.argv[2]) if len(sys.argv) > 2 else 32), 10
dev = torch.device("cuda:0")
k = Kernels(dev)
ens = PersistentEnsemble(d, dev)
g = torch.Generator(device=dev).manual_seed(20261018)
betas = [0.0] * 3 + list(np.geomspace(1e-4, 1.0, T - 3))
for t in range(T):
    u = torch.rand((n_gen, d), dtype=torch.float64, device=dev, generator=g)
    chi = (torch.randn((n_gen, d), dtype=torch.float64, device=dev, generator=g) ** 2).sum(1)
    ens.append(u, -0.5 * chi * 4.0, betas[t], -0.3 * t)
for i in range(6):
    k.probe(ens, 0.3 + 0.02 * i)
k.next_beta(ens, 0.2, 2.0 * n_gen, 0)
w = torch.empty(ens.n_total, dtype=torch.float64, device=dev)
k.weights(ens, 0.37, k.probe_out, w)
torch.cuda.synchronize()
print("ok", ens.n_total)
