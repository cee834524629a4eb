"""torchrun check: a sharded run must agree with a single-GPU run of the same problem (same
Philox seed) up to fp64 reduction order.  Run:  torchrun --nproc-per-node 2 tools/dist_check.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import tempest_b200 as tp

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = 10
def make():
    return tp.Sampler(tp.UniformPrior(-10.0, 10.0, d), tp.Rosenbrock(d), d, n_particles=n, vectorize=True,
                      clustering=False, random_state=123)
# single-GPU reference on every rank (before the process group exists)
s1 = make(); t0 = time.perf_counter(); s1.run(n_total=2048, progress=False); torch.cuda.synchronize(); t1 = time.perf_counter() - t0
b1, z1, st1 = s1.state.get_history("beta"), s1.evidence()[0], s1.state.get_history("steps")
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
s2 = make(); t0 = time.perf_counter(); s2.run(n_total=2048, progress=False); torch.cuda.synchronize(); t2 = time.perf_counter() - t0
b2, z2, st2 = s2.state.get_history("beta"), s2.evidence()[0], s2.state.get_history("steps")
x, w, l = s2.posterior()
if dist.get_rank() == 0:
    print("T", len(b1), len(b2), "logz", z1, z2, "time 1gpu %.3f  sharded %.3f" % (t1, t2))
    print("max |dbeta|", np.max(np.abs(b1 - b2)) if len(b1) == len(b2) else "len differs", "steps equal", np.array_equal(st1, st2))
    print("posterior rows", x.shape, "sum w", w.sum())
    assert len(b1) == len(b2) and np.allclose(b1, b2, rtol=1e-9, atol=1e-12) and abs(z1 - z2) < 1e-8 * abs(z1)
    print("DIST OK")
dist.destroy_process_group()
