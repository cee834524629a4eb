import os, sys, torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(1024, dtype=torch.float64, device=torch.device("cuda", local))
    h = symm.rendezvous(t, dist.group.WORLD)
    print("rank", dist.get_rank(), "buffer_ptrs", [hex(p) for p in h.buffer_ptrs], "signal", [hex(p) for p in h.signal_pad_ptrs][:2], flush=True)
    t.fill_(float(dist.get_rank() + 1)); torch.cuda.synchronize(); dist.barrier()
    peer = h.get_buffer((dist.get_rank() + 1) % dist.get_world_size(), (1024,), torch.float64)
    print("rank", dist.get_rank(), "peer value", float(peer[0].item()), flush=True)
except Exception as e:
    print("SYMM FAILED", type(e).__name__, e, flush=True)
dist.barrier(); dist.destroy_process_group()
