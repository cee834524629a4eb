"""Why is the first trained iteration (iteration 4) of a FRESH Sampler slow?  cProfile of exactly that iteration after a
warm-up Sampler has run and been dropped.  python tools/first_iter_profile.py  (or under torchrun)"""
import cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import tempest_b200 as tp

n, d = 1 << 20, 10
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank = dist.get_rank() if world > 1 else 0
make = lambda: tp.Sampler(tp.UniformPrior(-10.0, 10.0, d), tp.Rosenbrock(d), d, n_particles=n, vectorize=True,
                          clustering=False, random_state=20261018)
s = make(); s.run(progress=False); _ = s.posterior(); del s, _
for rep in range(2):
    s = make()
    core = s._core
    core._initialize_fresh(); core.n_total = 4096
    for _ in range(3):
        core.execute_iteration(export=False)
    torch.cuda.synchronize()
    before = torch.cuda.memory_stats()["num_device_alloc"] if hasattr(torch.cuda, "memory_stats") else 0
    pr = cProfile.Profile(); t0 = time.perf_counter(); pr.enable()
    core.execute_iteration(export=False)
    torch.cuda.synchronize()
    pr.disable(); dt = time.perf_counter() - t0
    after = torch.cuda.memory_stats()["num_device_alloc"]
    t1 = time.perf_counter(); core.execute_iteration(export=False); torch.cuda.synchronize(); dt5 = time.perf_counter() - t1
    if rank == 0:
        print(f"rep {rep}: iteration 4 took {dt * 1e3:.1f} ms (iteration 5: {dt5 * 1e3:.1f} ms); cudaMalloc calls during it: {after - before}")
        out = io.StringIO(); pstats.Stats(pr, stream=out).sort_stats("tottime").print_stats(14); print(out.getvalue()[:3500])
    del s, core
