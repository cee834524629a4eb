"""Micro-benchmarks of the non-MCMC kernels on a synthetic ensemble (CUDA events, L2-exceeding inputs)."""
import os, sys, json, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from tempest_b200 import _lib
from tempest_b200.ensemble import ptr, stream_ptr
from tempest_b200.steps import Kernels

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 25
d = 10
dev = torch.device("cuda:0")
lib = _lib.load()
k = Kernels(dev)
g = torch.Generator(device=dev); g.manual_seed(1)
u = torch.rand((n, d), dtype=torch.float64, device=dev, generator=g)
w = torch.rand(n, dtype=torch.float64, device=dev, generator=g) ** 8
w /= w.sum()
st = stream_ptr()
mws = k.ws.bytes("mom", lib.tb_moments_workspace_bytes(d))
mean = torch.zeros(d, dtype=torch.float64, device=dev); cov = torch.zeros(d * d, dtype=torch.float64, device=dev)
inv = torch.eye(d, dtype=torch.float64, device=dev).reshape(-1).contiguous(); out2 = torch.zeros(2, dtype=torch.float64, device=dev)
cdf = torch.empty(n, dtype=torch.float64, device=dev)
cws = k.ws.bytes("cdf_ws", lib.tb_cdf_workspace_bytes(n))
draws = torch.rand(1 << 20, dtype=torch.float64, device=dev, generator=g)
idx = torch.empty(1 << 20, dtype=torch.int64, device=dev)
cnt = torch.zeros(2048, dtype=torch.int64, device=dev); s1 = torch.zeros(2048, dtype=torch.float64, device=dev); s2 = torch.zeros_like(s1)
rws = k._reduce_ws
o3 = torch.zeros(3, dtype=torch.float64, device=dev)

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

res = {}
def rec(name, ms, nbytes): res[name] = dict(ms=round(ms, 4), gbs=round(nbytes / ms / 1e6, 1))
rec("mom_mean", timeit(lambda: lib.tb_moments_partial(ptr(u), None, ptr(w), None, n, d, 1.0, 1, 0, ptr(mws), ptr(mean), None, st)), n * 8 * (d + 1))
rec("mom_cov", timeit(lambda: lib.tb_moments_partial(ptr(u), None, ptr(w), None, n, d, 1.0, 0, 1, ptr(mws), ptr(mean), ptr(cov), st)), n * 8 * (d + 1))
rec("mahal_cv", timeit(lambda: lib.tb_mahalanobis_cv(ptr(u), ptr(w), n, d, ptr(mean), ptr(inv), ptr(rws), ptr(out2), st)), n * 8 * (d + 1))
rec("cdf_exact", timeit(lambda: lib.tb_cdf_exact(ptr(w), n, ptr(cdf), ptr(cws), st)), n * 16)
rec("search_right_1M", timeit(lambda: lib.tb_search_right(ptr(cdf), n, ptr(draws), 1 << 20, ptr(idx), st)), (1 << 20) * 16)
rec("binade_hist", timeit(lambda: lib.tb_binade_hist(ptr(w), n, ptr(cnt), ptr(s1), ptr(s2), st)), n * 8)
rec("masked_sums", timeit(lambda: lib.tb_masked_sums(ptr(w), n, 1e-9, ptr(rws), ptr(o3), st)), n * 8)
print(json.dumps(dict(n=n, d=d, **res)))
