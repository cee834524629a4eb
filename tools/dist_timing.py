"""Where does a sharded run spend host time?  torchrun --nproc-per-node G tools/dist_timing.py [n]
Times Sampler construction (symmetric-memory rendezvous), then whole runs without per-stage synchronisation, with the
wall clock of every iteration (rank 0)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import tempest_b200 as tp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
d = 10
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank = dist.get_rank() if world > 1 else 0


def say(*a):
    if rank == 0:
        print(*a, flush=True)


def sync():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for rep in range(3):
    sync()
    t0 = time.perf_counter()
    s = tp.Sampler(tp.UniformPrior(-10.0, 10.0, d), tp.Rosenbrock(d), d, n_particles=n, vectorize=True, clustering=False,
                   random_state=20261018)
    torch.cuda.synchronize()
    t_ctor = time.perf_counter() - t0
    core = s._core
    core._initialize_fresh()
    core.n_total = 4096
    marks = [time.perf_counter()]
    while core._not_termination():
        core.execute_iteration(export=False)
        marks.append(time.perf_counter())
    torch.cuda.synchronize()
    t_run = time.perf_counter() - marks[0]
    its = [round((b - a) * 1e3, 2) for a, b in zip(marks[:-1], marks[1:])]
    say(f"rep {rep}: constructor {t_ctor * 1e3:.1f} ms, run {t_run * 1e3:.1f} ms (T={len(its)}), host ms per iteration: {its}")
    t0 = time.perf_counter()
    x, w, l = s.posterior()
    say(f"   posterior() {1e3 * (time.perf_counter() - t0):.1f} ms, {x.shape[0]} samples; evidence {s.evidence()[0]:.6f}")
    del s, core, x, w, l
if world > 1:
    dist.destroy_process_group()
