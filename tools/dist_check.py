"""torchrun check of the sharded path against the single-GPU path of the same problem (same Philox seed).

    torchrun --nproc-per-node G tools/dist_check.py [n_particles] [--big]

1. Kernel level: the sharded exact cdf + two-level search must return numpy's indices on the GLOBAL weight vector,
   bit for bit (multinomial and systematic), for several weight distributions.
2. Run level: beta ladder, step counts, logZ of a sharded run vs the single-GPU run.  Philox counters are keyed by
   global walker slot and every discrete decision is taken from replicated quantities, but fp64 reductions are
   folded per rank first, so the contract is "independent of G up to fp64 reduction order": the check reports the
   first iteration whose discrete output differs and the size of the continuous differences.
3. Sharded features: systematic resampling, posterior(resample=True), checkpoint / resume.
Prints DIST OK when every assertion holds."""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import tempest_b200 as tp  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
args = [a for a in sys.argv[1:] if not a.startswith("--")]
n = int(args[0]) if args else 4096
big = "--big" in sys.argv
d = 10


def make(**kw):
    base = dict(n_particles=n, vectorize=True, clustering=False, random_state=123)
    base.update(kw)
    return tp.Sampler(tp.UniformPrior(-10.0, 10.0, d), tp.Rosenbrock(d), d, **base)


# ---- single-GPU reference on every rank (before the process group exists) -------------------------------------
s1 = make()
t0 = time.perf_counter()
s1.run(n_total=2048, progress=False)
torch.cuda.synchronize()
t1 = time.perf_counter() - t0
ref = {k: s1.state.get_history(k) for k in ("beta", "steps", "logz", "ess", "acceptance")}
z1 = s1.evidence()[0]
post1 = None if big else s1.posterior()
del s1
torch.cuda.empty_cache()

os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()


def say(*a):
    if rank == 0:
        print(*a, flush=True)


# ---- 1. kernel level ------------------------------------------------------------------------------------------------
from tempest_b200.sharded import ShardedKernels  # noqa: E402
from tempest_b200.dist import Comm  # noqa: E402
from oracle import ps_oracle as po  # noqa: E402

comm = Comm()
dev = torch.device("cuda", local)
k = ShardedKernels(dev, comm)
assert k.xgpu is not None, "peer-mapped memory unavailable"
rng = np.random.default_rng(5)
for kind, T, L in [("skewed", 7, 3000), ("zeros", 5, 1024), ("spiky", 3, 5000), ("equal", 4, 777), ("ladder", 12, 4096)]:
    NG = T * L * world                                     # global vector: T generations of world*L slots
    if kind == "skewed":
        p = np.exp(-0.5 * rng.chisquare(10, NG) * 40.0)
    elif kind == "zeros":
        p = rng.random(NG) * (rng.random(NG) < 0.3)
        p[: 2 * L] = 0.0
    elif kind == "spiky":
        p = rng.random(NG) * 1e-12
        p[rng.integers(0, NG, NG // 1000)] = 1.0
    elif kind == "equal":
        p = np.full(NG, 1.0 / NG)
    else:                                                  # weights growing by e^40 per generation: many binades
        p = np.exp(rng.normal(size=NG) * 3.0 + 40.0 * (np.arange(NG) // (L * world)))
    p = p / p.sum()
    pg = torch.as_tensor(p).to(dev)
    dist.broadcast(pg, 0)
    p = pg.cpu().numpy()
    # rank r stores slots [r L, (r+1) L) of every generation
    mine = np.concatenate([p[t * L * world + rank * L: t * L * world + (rank + 1) * L] for t in range(T)])
    pl = torch.as_tensor(mine).to(dev)
    seg = torch.arange(0, T + 1, dtype=torch.int64, device=dev) * L
    m = 4096
    ug = torch.as_tensor(rng.random(m)).to(dev)
    dist.broadcast(ug, 0)
    h = k.cdf_x(pl, T * L, seg, NG, "chk")
    idx = torch.empty(m, dtype=torch.int64, device=dev)
    k.search_x(h, ug, m, idx)
    cdf_ref = np.cumsum(p)
    loc_ref = np.concatenate([cdf_ref[t * L * world + rank * L: t * L * world + (rank + 1) * L] for t in range(T)])
    got = h["cdf"][: T * L].cpu().numpy()
    assert np.array_equal(got.view(np.uint64), loc_ref.view(np.uint64)), f"{kind}: sharded cdf differs from numpy cumsum"
    gidx = po.legacy_choice_indices(p, ug.cpu().numpy())              # global indices
    t_of, r_of, j_of = gidx // (L * world), (gidx % (L * world)) // L, gidx % L
    want = np.where(r_of == rank, t_of * L + j_of, -1)
    assert np.array_equal(idx.cpu().numpy(), want), f"{kind}: multinomial indices differ"
    u0 = 0.37
    k.search_x(h, None, m, idx, systematic=True, u0=u0)
    gidx = po.systematic_indices(m, p, u0)
    t_of, r_of, j_of = gidx // (L * world), (gidx % (L * world)) // L, gidx % L
    want = np.where(r_of == rank, t_of * L + j_of, -1)
    assert np.array_equal(idx.cpu().numpy(), want), f"{kind}: systematic indices differ"
    say(f"kernel check {kind}: cdf bitwise, multinomial + systematic indices equal; status {k.last_cdf_status[:5]}")

# ---- 2. run level -------------------------------------------------------------------------------------------------
s2 = make()
t0 = time.perf_counter()
s2.run(n_total=2048, progress=False)
torch.cuda.synchronize()
t2 = time.perf_counter() - t0
got = {kk: s2.state.get_history(kk) for kk in ref}
z2 = s2.evidence()[0]
T1, T2 = len(ref["beta"]), len(got["beta"])
Tm = min(T1, T2)
db = np.abs(ref["beta"][:Tm] - got["beta"][:Tm])
first_beta = int(np.argmax(db > 0)) if (db > 0).any() else None
st_eq = ref["steps"][:Tm] == got["steps"][:Tm]
first_steps = int(np.argmin(st_eq)) if not st_eq.all() else None
say(f"run check N={n}: T {T1} vs {T2}; logZ {z1!r} vs {z2!r} (|d| {abs(z1 - z2):.3e}); time 1 GPU {t1:.3f} s, {world} GPUs {t2:.3f} s")
say(f"   first iteration with a different beta: {first_beta}"
    + (f" (|d beta| / beta = {db[first_beta] / ref['beta'][first_beta]:.3e})" if first_beta is not None else "")
    + f"; first iteration with a different step count: {first_steps}; max |d beta| / beta over the run "
    f"{np.max(db[3:] / ref['beta'][3:Tm]) if Tm > 3 else 0.0:.3e}; max |d logZ_t| {np.max(np.abs(ref['logz'][:Tm] - got['logz'][:Tm])):.3e}")
# contract: same ladder up to fp64 reduction order; statistically equivalent evidence
assert abs(T1 - T2) <= 1
assert np.allclose(ref["beta"][:Tm], got["beta"][:Tm], rtol=2e-2 if big else 1e-6, atol=1e-12)
assert abs(z1 - z2) < (0.02 if big else 1e-6 * abs(z1) + 1e-6)
if not big:
    x, w, l = s2.posterior()
    assert x.shape == post1[0].shape and abs(w.sum() - 1.0) < 1e-9
    np.testing.assert_allclose(np.average(x, weights=w, axis=0), np.average(post1[0], weights=post1[1], axis=0), rtol=1e-6,
                               atol=1e-8)
    # ---- 3. sharded features ----------------------------------------------------------------------------------
    xr, wr, lr = s2.posterior(resample=True)
    assert xr.shape[0] == wr.shape[0] == lr.shape[0] and np.all(wr == 1.0 / len(wr))
    pm, pr = np.average(x, weights=w, axis=0), xr.mean(axis=0)
    assert np.all(np.abs(pm - pr) < 6.0 * x.std(axis=0) / np.sqrt(1.0 / np.sum(w * w))), "resampled posterior mean off"
    s3 = make(resample="syst")
    s3.run(n_total=2048, progress=False)
    z3 = s3.evidence()[0]
    say(f"   systematic resampling run: T {s3.state.get_history_length()}, logZ {z3:.4f}")
    assert abs(z3 - z2) < 1.0 and s3.beta == 1.0
    # checkpoint / resume: a resumed sharded run reproduces the uninterrupted one
    tmp = tempfile.gettempdir()
    sa = make(output_dir=os.path.join(tmp, "tb_dist_ckpt"), output_label="a")
    sa._core._initialize_fresh()
    sa._core.n_total = 2048
    for _ in range(6):
        sa._core.execute_iteration(export=False)
    path = os.path.join(tmp, "tb_dist_ckpt", "mid.state")
    sa._core.save_sampler_state(path)
    for _ in range(4):
        sa._core.execute_iteration(export=False)
    sb = make()
    sb._core.load_sampler_state(path)
    sb._core.n_total = 2048
    for _ in range(4):
        sb._core.execute_iteration(export=False)
    assert np.array_equal(sa.state.get_history("beta"), sb.state.get_history("beta"))
    assert np.array_equal(sa.state.get_history("logl"), sb.state.get_history("logl"))
    say("   posterior(resample=True), systematic resampling, checkpoint/resume: ok")
say("DIST OK")
dist.barrier()
dist.destroy_process_group()
