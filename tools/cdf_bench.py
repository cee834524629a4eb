"""Time tb_cdf_exact on the weight vector of a finished C4 run (run under ncu for the per-kernel split)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import tempest_b200 as tp
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
d = 10
s = tp.Sampler(tp.UniformPrior(-10.0, 10.0, d), tp.Rosenbrock(d), d, n_particles=n, vectorize=True, clustering=False,
               random_state=20261018)
s.run(progress=False)
core = s._core
ens, k = core.ensemble, core.k
k.probe(ens, 1.0)
w = k.weights(ens, 1.0, k.probe_out, core.weights_buffer())
k.g_normalize(w, ens.n_total)
for _ in range(3):
    k.cdf(w, ens.n_total)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    cdf = k.cdf(w, ens.n_total)
b.record()
torch.cuda.synchronize()
torch.cuda.profiler.start()
cdf = k.cdf(w, ens.n_total)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
ws = k.ws.bytes("cdf_ws", 64)
print("n", ens.n_total, "cdf ms", a.elapsed_time(b) / 10, "status", ws[:64].view(torch.int32).cpu().numpy()[:14])
import numpy as np
ref = np.cumsum(w.cpu().numpy())
print("bitwise equal to numpy:", np.array_equal(ref.view(np.uint64), cdf.cpu().numpy().view(np.uint64)))
