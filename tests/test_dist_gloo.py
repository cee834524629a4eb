"""Host-side logic of the sharded (multi-GPU) path, exercised with world_size = 2 on the gloo
backend (CPU): shard bounds, rank-order merge of ESS triples, the all-to-all that moves resampled
rows to their slot owners, and the padded variable-length all-gather."""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tempest_b200.dist import Comm, exchange_owned_rows, merge_ess_triples, shard_bounds


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = Comm()
        assert comm.on and comm.world == world and comm.rank == rank
        n, d = 64, 3
        lo, hi = shard_bounds(n, world, rank)
        # every rank knows the global table; each "owns" the slots whose ancestor it holds (odd/even here)
        rng = np.random.default_rng(0)
        table = torch.as_tensor(rng.random((n, d + 1)))
        owned = torch.arange(rank, n, world)
        mine = exchange_owned_rows(comm, table[owned].clone(), owned, n, lo, hi)
        assert torch.equal(mine, table[lo:hi])
        # uneven ownership (rank 0 holds the ancestors of 3/4 of the slots), incl. a rank that sends nothing to a peer
        cut = (3 * n) // 4
        owned = torch.arange(0, cut) if rank == 0 else torch.arange(cut, n)
        mine = exchange_owned_rows(comm, table[owned].clone(), owned, n, lo, hi)
        assert torch.equal(mine, table[lo:hi])
        # padded all-gather of variable-length shards, rank-major
        part = torch.arange(rank * 10, rank * 10 + 3 + rank, dtype=torch.float64).reshape(-1, 1)
        allp = comm.allgather_rows(part)
        want = torch.cat([torch.arange(r * 10, r * 10 + 3 + r, dtype=torch.float64) for r in range(world)])
        assert torch.equal(allp.flatten(), want)
        # probe triples: every rank merges the same list in rank order -> identical bits everywhere
        a = torch.as_tensor(rng.normal(size=1000) * 30.0)
        loc = a[rank::world]
        m = float(loc.max())
        w = torch.exp(loc - m)
        trip = torch.tensor([m, float(w.sum()), float((w * w).sum())], dtype=torch.float64)
        allt = comm.allgather(trip).numpy()
        mm, s1, s2 = merge_ess_triples(allt)
        wf = torch.exp(a - a.max())
        assert mm == float(a.max())
        assert s1 == pytest.approx(float(wf.sum()), rel=1e-13)
        assert s1 * s1 / s2 == pytest.approx(float(wf.sum() ** 2 / (wf * wf).sum()), rel=1e-12)
        tsum = comm.allreduce_sum_(torch.tensor([mm, s1, s2], dtype=torch.float64))
        assert tsum[0].item() == world * mm and tsum[1].item() == world * s1     # bitwise equal on all ranks
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_sharded_host_logic_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert dict(ret) == {0: "ok", 1: "ok"}


def test_shard_bounds_and_merge_edge_cases():
    assert shard_bounds(1 << 20, 8, 3) == (3 << 17, 4 << 17)
    with pytest.raises(ValueError):
        shard_bounds(10, 4, 0)
    assert merge_ess_triples([(-math.inf, 0.0, 0.0), (2.0, 3.0, 1.5)]) == (2.0, 3.0, 1.5)
    m, s1, s2 = merge_ess_triples([(0.0, 1.0, 1.0), (0.0, 1.0, 1.0)])
    assert (m, s1, s2) == (0.0, 2.0, 2.0)
    assert Comm().world == 1 and not Comm().on
