"""cProfile of the host side of one sharded (or single-GPU) run after a warm-up run: where does the Python thread spend
its time?  torchrun --nproc-per-node G tools/host_profile.py [n]"""
import cProfile, io, os, pstats, sys, time
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


def make():
    return tp.Sampler(tp.UniformPrior(-10.0, 10.0, d), tp.Rosenbrock(d), d, n_particles=n, vectorize=True, clustering=False,
                      random_state=20261018)


s = make()
s.run(progress=False)
del s
s = make()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
s.run(progress=False)
torch.cuda.synchronize()
pr.disable()
dt = time.perf_counter() - t0
if rank == 0:
    print(f"run {dt * 1e3:.1f} ms, T = {s.state.get_history_length()}")
    for key in ("cumulative", "tottime"):
        out = io.StringIO()
        pstats.Stats(pr, stream=out).sort_stats(key).print_stats(45)
        print(out.getvalue()[:9000])
if world > 1:
    dist.destroy_process_group()
