"""Per-iteration stage timings of one full run (CUDA events); writes gpurun_out/profile_run.json."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import tempest_b200 as tp  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
d = 10
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if local != 0:
        sys.stdout = open(os.devnull, "w")
for _warm in range(int(os.environ.get("PROFILE_WARM_RUNS", "0"))):     # untimed: module loading, allocations
    _s = tp.Sampler(tp.UniformPrior(-10.0, 10.0, d), tp.Rosenbrock(d), d, n_particles=n, vectorize=True,
                    clustering=False, random_state=20261018)
    _s.run(progress=False)
    del _s
s = tp.Sampler(tp.UniformPrior(-10.0, 10.0, d), tp.Rosenbrock(d), d, n_particles=n, vectorize=True,
               clustering=False, random_state=20261018)
core = s._core
core.profile = True
core._initialize_fresh()
core.n_total = 4096
rows = []
torch.cuda.synchronize()
t_all = time.perf_counter()
while core._not_termination():
    core.stage_ms = {}
    t0 = time.perf_counter()
    core.execute_iteration(export=False)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    st = core.state
    row = dict(iter=st.raw("iter"), beta=st.raw("beta"), steps=st.raw("steps"), n_hist=core.ensemble.n_total,
               wall_ms=wall, n_trim=getattr(core.k, "last_trim", {}).get("n_trim"),
               n_exact=getattr(core.k, "last_trim", {}).get("n_exact"),
               probes=len(core.reweighter.probe_log), **{k: round(v, 3) for k, v in core.stage_ms.items()})
    rows.append(row)
    print(json.dumps(row), flush=True)
total = time.perf_counter() - t_all
agg = {}
for r in rows[5:]:
    for k_, v in r.items():
        if isinstance(v, float) and k_ not in ("beta", "wall_ms"):
            agg[k_] = agg.get(k_, 0.0) + v
print("stage sums (iterations 6..T, ms):", {k_: round(v, 1) for k_, v in sorted(agg.items(), key=lambda kv: -kv[1])})
print(f"T={len(rows)} total {total:.3f} s  -> {len(rows) / total:.2f} it/s, calls {core.state.raw('calls')}, "
      f"{core.state.raw('calls') / total:.3e} evals/s")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
if local == 0:
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"profile_run_g{world}.json"), "w"))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
