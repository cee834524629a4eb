"""Latency of the in-kernel synchronisation point (csrc/tb_xgpu.cuh grid_xreduce): 10^4 back-to-back grid-wide
reductions -- and exchanges over NVLink peer memory when run under torchrun -- inside ONE kernel.

    python tools/xgpu_bench.py                      (one GPU: the grid barrier + fold alone)
    torchrun --nproc-per-node G tools/xgpu_bench.py (G GPUs: + the peer stores / flag waits)

grid = 1 CTA is the pure exchange; grid = 1036 CTAs (7 per SM, the Metropolis kernel's shape) and 592 (4 per SM, the ESS
search's) are what the persistent kernels pay per Metropolis step / per ESS pass."""
import ctypes as C
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from tempest_b200 import _lib
from tempest_b200.ensemble import ptr, stream_ptr

local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
lib = _lib.load()
xref = None
k = None
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    from tempest_b200.dist import Comm
    from tempest_b200.sharded import ShardedKernels

    k = ShardedKernels(dev, Comm())
    assert k.xgpu is not None, "peer memory unavailable"
rank = dist.get_rank() if world > 1 else 0
count = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
lines = [f"# tools/xgpu_bench.py: {count} back-to-back synchronisation points in one cooperative launch, {world} GPU(s)"]
sm = int(lib.tb_sm_count())
for grid in (1, sm, 4 * sm, 7 * sm):
    ws = torch.zeros(int(lib.tb_xgpu_bench_workspace_bytes(grid)) + 256, dtype=torch.uint8, device=dev)
    out = torch.zeros(2, dtype=torch.float64, device=dev)
    best = None
    for rep in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        rc = lib.tb_xgpu_bench(C.byref(k.xgpu) if k is not None else None, grid, count, ptr(ws), ptr(out), stream_ptr())
        assert rc == 0, rc
        torch.cuda.synchronize()
        if k is not None:
            k.consume_exchanges(count)
        ns = float(out[0].item())
        best = ns if best is None else min(best, ns)
    t = torch.tensor([best], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    expect = sum((i & 7) for i in range(count)) * grid * world
    ok = float(out[1].item()) == float(expect)
    lines.append(f"grid {grid:5d} CTAs x 128 threads: {t.item() / 1e3:7.2f} us per synchronisation point (max over ranks, best of 3); "
                 f"checksum {'ok' if ok else 'MISMATCH'}")
if rank == 0:
    print("\n".join(lines))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", f"xgpu_bench_{world}gpu.txt"), "w").write("\n".join(lines) + "\n")
if world > 1:
    dist.destroy_process_group()
