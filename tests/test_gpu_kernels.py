"""Kernel-level parity: every C-ABI entry point against the CPU oracle / numpy on the same inputs.

Bit-exact for indices, order statistics, cumulative sums and in-kernel likelihoods; 1e-10
relative (north-star tolerance) for fp64 reductions whose summation order differs.
"""
import ctypes as C
import math
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from conftest import GOLDEN  # noqa: E402

RTOL = 1e-10  # BASELINE.json north_star: "within 1e-10 relative in fp64"


@pytest.fixture(scope="module")
def env():
    from tempest_b200 import _lib
    from tempest_b200.ensemble import PersistentEnsemble, ptr, stream_ptr
    from tempest_b200.steps import Kernels

    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)

    class Env:
        pass

    e = Env()
    e.lib, e.dev, e.ptr, e.sp, e.k = _lib.load(), dev, ptr, stream_ptr, Kernels(dev)
    e.Ensemble = PersistentEnsemble
    e._lib = _lib
    return e


def dev_arr(env, a, dtype=None):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(env.dev)


def synthetic_ensemble(env, n, T, d, seed=0, tau=1.0):
    """SURVEY 8d synthetic persistent ensemble: logl ~ -0.5 chi2_d * tau, beta geometric after 3 zeros."""
    from oracle import ps_oracle as po

    rng = np.random.default_rng(seed)
    betas = [0.0, 0.0, 0.0] + list(np.geomspace(1e-4, 1.0, max(T - 3, 1)))[: max(T - 3, 0)]
    betas = betas[:T]
    ens = env.Ensemble(d, env.dev)
    gens_l, gens_u, logzs = [], [], []
    for t in range(T):
        u = rng.random((n, d))
        logl = -0.5 * tau * rng.chisquare(d, n) * (1.0 + 3.0 * rng.random(n))
        logz = 0.0 if t == 0 else po.log_weights_and_logz(gens_l, betas[:t], logzs, betas[t])[1]
        ens.append(dev_arr(env, u), dev_arr(env, logl), betas[t], logz)
        gens_l.append(logl)
        gens_u.append(u)
        logzs.append(logz)
    return ens, gens_l, gens_u, betas, logzs


# ---------------------------------------------------------------------------------------
def test_mixture_probe_weights_match_oracle(env):
    from oracle import ps_oracle as po

    for (n, T, d) in [(64, 5, 3), (1000, 9, 10), (4096, 24, 10)]:
        ens, gl, gu, betas, logzs = synthetic_ensemble(env, n, T, d, seed=n)
        c_inc = ens.C[: ens.n_total].clone()
        ens.rebuild_mixture()
        assert torch.equal(c_inc, ens.C[: ens.n_total]), "incremental log-mixture != full rebuild"
        for beta in (0.0, betas[-1], 0.37, 1.0):
            logw_ref, logz_ref = po.log_weights_and_logz(gl, betas, logzs, beta)
            w_ref = np.exp(logw_ref - logw_ref.max())
            ess_ref = po.effective_sample_size(w_ref)
            out = env.k.probe(ens, beta).cpu().numpy()
            assert out[5] == 0
            assert out[3] == pytest.approx(ess_ref, rel=RTOL)
            assert out[4] == pytest.approx(logz_ref, rel=RTOL, abs=1e-12)
            w = torch.empty(ens.n_total, dtype=torch.float64, device=env.dev)
            env.k.weights(ens, beta, env.k.probe_out, w)
            np.testing.assert_allclose(w.cpu().numpy(), w_ref / w_ref.sum(), rtol=RTOL, atol=1e-300)
            env.k.weights(ens, beta, env.k.probe_out, w, log=True)
            np.testing.assert_allclose(w.cpu().numpy(), logw_ref, rtol=RTOL, atol=1e-9)


def test_next_beta_device_search_matches_oracle_probe_sequence(env):
    from oracle import ps_oracle as po

    for (n, T, d, seed) in [(256, 6, 4, 1), (1024, 12, 10, 2), (5000, 20, 10, 3)]:
        ens, gl, gu, betas, logzs = synthetic_ensemble(env, n, T, d, seed=seed)
        beta_prev = betas[-1] * 0.5
        target = 2.0 * n

        def probe(b):
            lw, _ = po.log_weights_and_logz(gl, betas, logzs, b)
            w = np.exp(lw - lw.max())
            return w, po.effective_sample_size(w), 0.0

        s = po.BetaSearch(probe, False)
        lo, hi = s.ess_bracket(beta_prev, target)
        if lo == hi:
            beta_ref = lo
            _, ess_ref, _ = s.probe(lo)
        else:
            beta_ref, _, ess_ref = s.bisect(beta_prev, hi, target, False)
        # speculative passes (three betas per pass, flags bit 1): same probe sequence and result from about half the passes
        res2, plog2 = env.k.next_beta(ens, beta_prev, target, 2)      # (results live in reused workspace buffers: copy)
        h2 = res2.cpu().numpy().copy()
        log2 = plog2[: 2 * int(h2[6])].cpu().numpy().copy()
        res, plog = env.k.next_beta(ens, beta_prev, target, 0)
        h = res.cpu().numpy()
        assert h2[0] == h[0] and h2[6] == h[6] and h2[9] <= h[9] and h[9] == h[6]
        np.testing.assert_allclose(h2[1:6], h[1:6], rtol=1e-12)
        np.testing.assert_array_equal(log2[0::2], plog[: 2 * int(h[6])].cpu().numpy()[0::2])
        nprobe = int(h[6])
        got = plog[: 2 * nprobe].cpu().numpy().reshape(-1, 2)
        # the oracle re-probes beta when lo == hi; the device reuses the probe it already has
        ref_log = np.array(s.log if lo != hi else s.log[:-1])
        assert h[0] == beta_ref                                    # bit-exact beta
        np.testing.assert_array_equal(got[:, 0], ref_log[:, 0])    # identical probe sequence
        np.testing.assert_allclose(got[:, 1], ref_log[:, 1], rtol=RTOL)
        assert h[4] == pytest.approx(ess_ref, rel=RTOL)
        # the one-launch search equals the host-driven search kernel by kernel
        out = env.k.probe(ens, h[0]).cpu().numpy()
        np.testing.assert_array_equal(out[:5], h[1:6])


@pytest.mark.parametrize("kind", ["uniform", "skewed", "tiny", "zeros", "ascending", "equal", "spiky"])
@pytest.mark.parametrize("n", [1, 2, 31, 512, 513, 4097, 100003, 1 << 20])
def test_cdf_exact_is_bitwise_numpy_cumsum(env, kind, n):
    rng = np.random.default_rng(hash((kind, n)) % (2**32))
    if kind == "uniform":
        p = rng.random(n)
    elif kind == "skewed":
        p = np.exp(-0.5 * rng.chisquare(10, n) * 40.0)
    elif kind == "tiny":
        p = rng.random(n) * 1e-300
    elif kind == "zeros":
        p = rng.random(n) * (rng.random(n) < 0.3)
        p[: min(n, 5)] = 0.0
    elif kind == "ascending":
        p = np.sort(np.exp(rng.normal(size=n) * 30.0))
    elif kind == "equal":
        p = np.full(n, 1.0 / n)
    else:
        p = rng.random(n) * 1e-12
        p[rng.integers(0, n, size=max(1, n // 1000))] = 1.0
    if p.sum() > 0:
        p = p / p.sum()
    ref = np.cumsum(p)
    dp = dev_arr(env, p)
    out = env.k.cdf(dp, n, "t_cdf")
    got = out.cpu().numpy()
    assert np.array_equal(got.view(np.uint64), ref.view(np.uint64)), \
        f"first mismatch at {np.nonzero(got != ref)[0][:3]}"


def _cdf_bitwise(env, p, name="t_cdf"):
    ref = np.cumsum(p)
    got = env.k.cdf(dev_arr(env, p), len(p), name).cpu().numpy()
    assert np.array_equal(got.view(np.uint64), ref.view(np.uint64)), \
        f"first mismatch at {np.nonzero(got.view(np.uint64) != ref.view(np.uint64))[0][:3]} of {len(p)}"


@pytest.mark.parametrize("chain", [1, 0])
def test_cdf_exact_hard_cases_both_paths(env, chain):
    """Round-half ties, oversized elements, binade crossings on tile boundaries, long zero / subnormal stretches and a
    PS-shaped vector (generations of rising weight), through the chained look-back kernel (chain = 1) and the
    multi-kernel pipeline (chain = 0, the default and the path sharded runs use)."""
    env.lib.tb_cdf_set_chain(chain)
    try:
        rng = np.random.default_rng(77)
        cases = []
        # exact ties: s = 1, then elements of exactly half an ulp (round-half-even decides), mixed with ordinary ones
        t = np.full(5000, 2.0 ** -53)
        t[0] = 1.0
        t[rng.integers(1, 5000, 200)] = rng.random(200) * 1e-9
        cases.append(t)
        # every element dominates the running sum (each add crosses at least one binade)
        cases.append(3.0 ** np.arange(600) * 1e-200)
        # a crossing exactly at the tile boundaries of both kernels (512 / 1024 elements)
        b = np.full(4096, 2.0 ** -13)
        cases.append(b)                                   # sum hits powers of two at indices 2^k - 1
        # leading zeros over many tiles, then subnormals, then ordinary weights
        z = np.zeros(70000)
        z[30000:30500] = 5e-324 * rng.integers(1, 1000, 500)
        z[40000:] = rng.random(30000) * 1e-3
        z[50000] = 0.7
        cases.append(z)
        # tiny normal range (below the integer regime), slowly growing into it
        cases.append(np.exp(rng.uniform(-700.0, -640.0, 20000)))
        # PS-shaped: 24 generations of 8192, log-weights rising by ~12 per generation with heavy scatter, early ones 0
        g = np.concatenate([np.exp(np.minimum(0.0, -300.0 + 12.5 * t_ + 6.0 * rng.standard_normal(8192))) for t_ in range(24)])
        g[: 3 * 8192] = 0.0
        cases.append(g / g.sum())
        # large: 2^22 skewed (many tiles in flight behind the few crossings)
        big = np.exp(-0.5 * rng.chisquare(10, 1 << 22) * 30.0)
        cases.append(big / big.sum())
        for c in cases:
            _cdf_bitwise(env, c)
    finally:
        env.lib.tb_cdf_set_chain(0)


@pytest.mark.parametrize("chain", [0, 1])
def test_cdf_exact_large_and_real_weights(env, chain):
    """2^25 elements, and the weight vector of a finished PS run at beta = 1 (the vector the resampling step sees)."""
    import tempest_b200 as tp

    env.lib.tb_cdf_set_chain(chain)

    rng = np.random.default_rng(78)
    p = rng.random(1 << 25) ** 8
    p /= p.sum()
    _cdf_bitwise(env, p, "t_cdf_big")
    s = tp.Sampler(tp.UniformPrior(-10.0, 10.0, 10), tp.Rosenbrock(10), 10, n_particles=1 << 15, vectorize=True,
                   clustering=False, random_state=3)
    s.run(progress=False)
    core = s._core
    ens, k = core.ensemble, core.k
    k.probe(ens, 1.0)
    w = k.weights(ens, 1.0, k.probe_out, core.weights_buffer())
    k.g_normalize(w, ens.n_total)
    wh = w.cpu().numpy().copy()
    got = k.cdf(w, ens.n_total, "t_cdf_real").cpu().numpy()
    env.lib.tb_cdf_set_chain(0)
    assert np.array_equal(got.view(np.uint64), np.cumsum(wh).view(np.uint64))


def test_cdf_sequential_kernel_agrees(env):
    rng = np.random.default_rng(9)
    n = 50000
    p = rng.random(n) ** 3
    p /= p.sum()
    dp = dev_arr(env, p)
    out = torch.empty(n, dtype=torch.float64, device=env.dev)
    env._lib.check(env.lib.tb_cdf_sequential(env.ptr(dp), n, env.ptr(out), env.sp()))
    assert np.array_equal(out.cpu().numpy(), np.cumsum(p))


def test_multinomial_and_systematic_indices_bit_exact(env):
    from oracle import ps_oracle as po

    rng = np.random.default_rng(4)
    for n, m in [(100, 64), (5000, 4096), (1 << 18, 1 << 16)]:
        p = np.exp(-0.5 * rng.chisquare(6, n) * 5.0)
        p /= p.sum()
        u = rng.random(m)
        ref = po.legacy_choice_indices(p, u)
        dp = dev_arr(env, p)
        cdf = env.k.cdf(dp, n, "t_cdf")
        idx = torch.empty(m, dtype=torch.int64, device=env.dev)
        du = dev_arr(env, u)
        env.k.search_right(cdf, n, du, idx)
        np.testing.assert_array_equal(idx.cpu().numpy(), ref)
        u0 = float(rng.random())
        ref_s = po.systematic_indices(m, p, u0)
        env.k.systematic(cdf, n, u0, m, idx)
        np.testing.assert_array_equal(idx.cpu().numpy(), ref_s)
    # edge: draw exactly on a boundary / at 0
    p = np.array([0.25, 0.25, 0.5])
    cdf = env.k.cdf(dev_arr(env, p), 3, "t_cdf")
    u = np.array([0.0, 0.25, 0.5, 0.4999999999999999, 0.9999999999999999])
    idx = torch.empty(5, dtype=torch.int64, device=env.dev)
    du = dev_arr(env, u)
    env.k.search_right(cdf, 3, du, idx)
    np.testing.assert_array_equal(idx.cpu().numpy(), po.legacy_choice_indices(p, u))


def test_gather_rows(env):
    rng = np.random.default_rng(5)
    n, d, m = 1000, 10, 333
    u = rng.random((n, d))
    l = rng.normal(size=n)
    idx = rng.integers(0, n, m)
    au = torch.empty((m, d), dtype=torch.float64, device=env.dev)
    al = torch.empty(m, dtype=torch.float64, device=env.dev)
    du, dl, di = dev_arr(env, u), dev_arr(env, l), dev_arr(env, idx)   # keep the inputs alive
    env._lib.check(env.lib.tb_gather_rows(env.ptr(du), env.ptr(dl), d, env.ptr(di), m, env.ptr(au), env.ptr(al),
                                          env.sp()))
    np.testing.assert_array_equal(au.cpu().numpy(), u[idx])
    np.testing.assert_array_equal(al.cpu().numpy(), l[idx])


@pytest.mark.parametrize("d", [1, 2, 7, 10, 16, 50, 100])
def test_volume_variation_matches_oracle(env, d):
    from oracle import ps_oracle as po

    rng = np.random.default_rng(d)
    n = 4000 if d <= 16 else 1500
    u = rng.random((n, d))
    w = np.exp(-0.5 * rng.chisquare(4, n))
    w /= w.sum()
    du, dw = dev_arr(env, u), dev_arr(env, w)
    got = env.k.volume_variation(du, dw, n, d)
    assert got == pytest.approx(po.volume_variation(u, w), rel=1e-9)
    # too few samples -> 1e10 (tools.py:87-88)
    assert env.k.volume_variation(du, dw, d, d) == 1e10


def test_volume_variation_rank_deficient_regularises(env):
    from oracle import ps_oracle as po

    rng = np.random.default_rng(1)
    n, d = 500, 4
    u = rng.random((n, d))
    u[:, 3] = u[:, 0]                       # exactly collinear -> matrix_rank < d
    w = np.full(n, 1.0 / n)
    du, dw = dev_arr(env, u), dev_arr(env, w)
    got = env.k.volume_variation(du, dw, n, d)
    assert got == pytest.approx(po.volume_variation(u, w), rel=1e-6)


@pytest.mark.parametrize("n,kind", [(40, "flat"), (1000, "skewed"), (20000, "skewed"), (300000, "heavy"),
                                    (4096, "equal"), (5000, "dups")])
def test_trim_matches_oracle(env, n, kind):
    from oracle import ps_oracle as po

    rng = np.random.default_rng(n)
    if kind == "flat":
        w = 1.0 + 0.1 * rng.random(n)
    elif kind == "skewed":
        w = np.exp(-0.5 * rng.chisquare(10, n))
    elif kind == "heavy":
        w = np.exp(-0.5 * rng.chisquare(10, n) * 8.0)
    elif kind == "equal":
        w = np.ones(n)
    else:
        w = np.repeat(np.exp(-rng.chisquare(3, n // 10)), 10)
    w_ref = w.copy()
    idx_ref, wt_ref, i_ref = po.trim_weights(w_ref)
    dw = dev_arr(env, w)
    idx, wt = env.k.trim(dw, n)
    assert env.k.last_trim["bin"] == i_ref
    np.testing.assert_array_equal(idx.cpu().numpy(), idx_ref)
    np.testing.assert_allclose(wt.cpu().numpy(), wt_ref, rtol=RTOL)
    np.testing.assert_allclose(dw.cpu().numpy(), w_ref, rtol=RTOL)   # normalised in place (tools.py:36)


def test_select_ranks_exact(env):
    rng = np.random.default_rng(6)
    n, d = 3001, 5
    u = rng.random((n, d))
    u[::7] = u[3]                                    # ties
    rows = rng.permutation(n)[:2000].astype(np.int64)
    mult = rng.integers(0, 5, size=2000).astype(np.int32)
    m_total = int(mult.sum())
    ranks = np.array([m_total // 2 - 1, m_total // 2], dtype=np.int64)
    out = torch.empty(2 * d, dtype=torch.float64, device=env.dev)
    ws = torch.zeros(env.lib.tb_select_workspace_bytes(d, 2), dtype=torch.uint8, device=env.dev)
    du, dr, dm, dk = dev_arr(env, u), dev_arr(env, rows), dev_arr(env, mult), dev_arr(env, ranks)
    env._lib.check(env.lib.tb_select_ranks(env.ptr(du), env.ptr(dr), d, 2000, d, env.ptr(dm), env.ptr(dk), 2,
                                           env.ptr(ws), env.ptr(out), env.sp()))
    expanded = np.repeat(u[rows], mult, axis=0)
    srt = np.sort(expanded, axis=0)
    np.testing.assert_array_equal(out.cpu().numpy().reshape(d, 2), srt[ranks].T)


@pytest.mark.parametrize("d", [1, 3, 10, 50, 100])
def test_chol_inv_and_regularisation(env, d):
    rng = np.random.default_rng(d)
    a = rng.normal(size=(d, d))
    spd = a @ a.T / d + 0.1 * np.eye(d)
    bad = spd.copy()
    bad[:, -1] = 0.0
    bad[-1, :] = 0.0                                 # exact zero pivot -> LinAlgError -> regularised (modes.py:114-119)
    mats = np.stack([spd, bad]) if d > 1 else np.stack([spd, np.zeros((1, 1))])
    dm = dev_arr(env, mats)
    chol = torch.empty_like(dm)
    inv = torch.empty_like(dm)
    info = torch.zeros(2, dtype=torch.int32, device=env.dev)
    env._lib.check(env.lib.tb_chol_inv(env.ptr(dm), d, 2, env.ptr(chol), env.ptr(inv), env.ptr(info), None, env.sp()))
    assert info.cpu().tolist() == [0, 1]
    np.testing.assert_allclose(chol[0].cpu().numpy(), np.linalg.cholesky(spd), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(inv[0].cpu().numpy(), np.linalg.inv(spd), rtol=1e-8, atol=1e-10)
    reg = mats[1] + np.eye(d) * max(1e-6, 1e-6 * abs(np.trace(mats[1])))
    np.testing.assert_allclose(dm[1].cpu().numpy(), reg, rtol=1e-14)
    np.testing.assert_allclose(chol[1].cpu().numpy(), np.linalg.cholesky(reg), rtol=1e-6, atol=1e-9)


def _params(env, prior, like, d, rng_mode=0, seed=7, iteration=1, **kw):
    p = env._lib.TbMcmcParams()
    p.n_dim, p.n_modes, p.sampler, p.rng_mode = d, 1, 0, rng_mode
    p.like_id, p.prior_id, p.n_steps, p.n_max = like.kernel_id, prior.kernel_id, 1, 20
    p.beta, p.seed, p.iteration, p.slot_offset, p.n_global = 0.0, seed, iteration, 0, 0
    keep = [dev_arr(env, like.dparams()), dev_arr(env, prior.dparams())]
    p.like_params, p.prior_params = keep[0].data_ptr(), keep[1].data_ptr()
    return p, keep


def test_prior_draw_and_likelihoods_bit_exact(env):
    from tempest_b200 import registry as reg

    rng = np.random.default_rng(8)
    cases = [
        (reg.UniformPrior(-10, 10, 10), reg.Rosenbrock(10)),
        (reg.UniformPrior(-10, 10, 50), reg.GaussianLikelihood.ar1(50)),
        (reg.UniformPrior(-10, 10, 2), reg.IsotropicMixture.four_corners(2)),
        (reg.UniformPrior(-6, 6, 100), reg.TwinShells(100)),
        (reg.UniformPrior(-6, 6, 4), reg.TwinShells(4)),
    ]
    for prior, like in cases:
        d, n = prior.n_dim, 777
        tape_u = rng.random((n, d))
        p, keep = _params(env, prior, like, d, rng_mode=1)
        u = torch.empty((n, d), dtype=torch.float64, device=env.dev)
        x = torch.empty((n, d), dtype=torch.float64, device=env.dev)
        logl = torch.empty(n, dtype=torch.float64, device=env.dev)
        dt = dev_arr(env, tape_u)
        env._lib.check(env.lib.tb_prior_draw(n, C.byref(p), env.ptr(dt), env.ptr(u), env.ptr(x), env.ptr(logl),
                                             env.sp()))
        np.testing.assert_array_equal(u.cpu().numpy(), tape_u)
        x_ref = np.array([prior(r) for r in tape_u])
        np.testing.assert_array_equal(x.cpu().numpy(), x_ref)
        ref = like(x_ref)
        got = logl.cpu().numpy()
        if like.kernel_id in (reg.LIKE_ROSENBROCK, reg.LIKE_GAUSSIAN):
            np.testing.assert_array_equal(got, ref)               # only + - * : bit-exact
        else:
            np.testing.assert_allclose(got, ref, rtol=1e-13)       # exp/log/sqrt differ by ulps


def test_philox_uniforms_are_reproducible_and_uniform(env):
    n = 1 << 16
    a = torch.empty(n, dtype=torch.float64, device=env.dev)
    b = torch.empty(n, dtype=torch.float64, device=env.dev)
    env._lib.check(env.lib.tb_philox_uniform(123, 5, 4, 0, n, env.ptr(a), env.sp()))
    env._lib.check(env.lib.tb_philox_uniform(123, 5, 4, 0, n, env.ptr(b), env.sp()))
    assert torch.equal(a, b)
    env._lib.check(env.lib.tb_philox_uniform(123, 6, 4, 0, n, env.ptr(b), env.sp()))
    assert not torch.equal(a, b)
    h = a.cpu().numpy()
    assert 0.0 <= h.min() and h.max() < 1.0
    assert abs(h.mean() - 0.5) < 0.01 and abs(h.var() - 1 / 12) < 0.005
    # sharding invariance: offset selects a window of the same stream
    env._lib.check(env.lib.tb_philox_uniform(123, 5, 4, 1000, 100, env.ptr(b), env.sp()))
    assert torch.equal(a[1000:1100], b[:100])


def _variates(env, n, d, shape, family, step=3, attempt=2, seed=987654321, iteration=9, slot_offset=0,
              want=("gamma", "z", "acc")):
    g = torch.empty(n, dtype=torch.float64, device=env.dev) if "gamma" in want else None
    z = torch.empty((n, d), dtype=torch.float64, device=env.dev) if "z" in want else None
    a = torch.empty(n, dtype=torch.float64, device=env.dev) if "acc" in want else None
    env._lib.check(env.lib.tb_debug_variates(seed, iteration, slot_offset, n, step, d, float(shape), attempt, family,
                                             env.ptr(g), env.ptr(z), env.ptr(a), env.sp()), "tb_debug_variates")
    return g, z, a


@pytest.mark.parametrize("shape", [0.5 * (10 + 1e6), 1.5])
def test_production_variates_match_numpy_restatement(env, shape):
    """The variates a Philox-mode Metropolis step consumes (the code tape mode replaces: tb::gamma_mt, normals_fixed,
    accept_uniform) against oracle/philox.py, walker by walker.  The device evaluates the Box-Muller radius / angle
    with MUFU log / sin / cos (abs. error ~2^-21), so equality is to 2e-5 absolute on |z| <= 6.76."""
    from oracle import philox as ph

    n, d, step, attempt, seed, it, off = 1 << 16, 10, 3, 2, 987654321, 9, 5_000_000_000
    g, z, a = _variates(env, n, d, shape, 0, step, attempt, seed, it, off)
    slots = off + np.arange(n, dtype=np.uint64)
    zr = ph.step_normals(seed, it, slots, step, attempt, d)
    dz = np.abs(z.cpu().numpy() - zr)
    assert dz.max() < 2e-4 and np.quantile(dz, 0.9999) < 1e-5
    gr, ar, margin = ph.step_gamma(seed, it, slots, step, shape)
    gd, ad = g.cpu().numpy(), a.cpu().numpy()
    # a Marsaglia-Tsang comparison within 1e-5 of its threshold may resolve differently (fp32 normal inside it)
    safe = margin > 1e-5
    assert safe.mean() > 0.999
    np.testing.assert_allclose(gd[safe], gr[safe], rtol=1e-7 if shape > 1e3 else 2e-4)   # (1 + t)^3 with a MUFU-rounded normal
    np.testing.assert_array_equal(ad[safe], ar[safe])             # the accept uniform is exact (one 32-bit word)
    assert 0.0 < ad.min() and ad.max() < 1.0
    # sharding invariance: the draws depend on the global slot only
    g2, z2, a2 = _variates(env, 100, d, shape, 0, step, attempt, seed, it, off + 4000)
    assert torch.equal(z2, z[4000:4100]) and torch.equal(g2, g[4000:4100]) and torch.equal(a2, a[4000:4100])


def _ks_statistic(x: "torch.Tensor", cdf) -> float:
    xs, _ = torch.sort(x.reshape(-1))
    n = xs.numel()
    f = cdf(xs)
    i = torch.arange(1, n + 1, device=xs.device, dtype=torch.float64)
    return float(torch.maximum((i / n - f).max(), (f - (i - 1) / n).max()))


@pytest.mark.parametrize("family", [0, 1])
def test_production_variates_distribution(env, family):
    """Kolmogorov-Smirnov and moment checks of the production normals (10^8 draws for the fused kernels' fp32
    Box-Muller), the Marsaglia-Tsang gamma (large shape through the series branch, small shape through the log
    branch) and the accept uniform.  KS bound 1.95 / sqrt(n) is the 0.1 % critical value."""
    from scipy import stats

    n = (10_000_000 if family == 0 else 2_000_000)
    d = 10
    _, z, _ = _variates(env, n, d, 2.0, family, want=("z",))
    m = n * d
    assert _ks_statistic(z, lambda t: torch.special.ndtr(t)) < 1.95 / math.sqrt(m)
    mean, var = float(z.mean()), float(z.var())
    assert abs(mean) < 5.0 / math.sqrt(m) and abs(var - 1.0) < 5.0 * math.sqrt(2.0 / m)
    assert abs(float((z ** 4).mean()) - 3.0) < 5.0 * math.sqrt(96.0 / m)
    # independence across coordinates of a walker (blocks of four normals come from one Philox block)
    zc = z[:1_000_000]
    corr = torch.corrcoef(zc.T) - torch.eye(d, dtype=torch.float64, device=env.dev)
    assert float(corr.abs().max()) < 6.0 / math.sqrt(zc.shape[0])
    del z, zc
    for shape in (0.5 * (10 + 1e6), 26.0, 1.0):
        ng = 4_000_000
        g, _, a = _variates(env, ng, d, shape, family, want=("gamma", "acc"))
        gh, ah = g.cpu().numpy(), a.cpu().numpy()
        ks = stats.kstest(gh, "gamma", args=(shape,)).statistic
        assert ks < 1.95 / math.sqrt(ng), (shape, ks)
        assert abs(gh.mean() / shape - 1.0) < 5.0 / math.sqrt(ng * shape)
        assert abs(gh.var() / shape - 1.0) < 5.0 * math.sqrt(2.0 / ng) + 5.0 * math.sqrt(6.0 / (shape * ng))
        assert stats.kstest(ah, "uniform").statistic < 1.95 / math.sqrt(ng)
        # the accept uniform shares a Philox block with the accepted gamma trial: must be uncorrelated with it
        assert abs(np.corrcoef(gh, ah)[0, 1]) < 5.0 / math.sqrt(ng)


@pytest.mark.parametrize("n,d", [(1, 1), (2, 3), (4097, 1), (3001, 5), (200000, 10)])
def test_select_pair_exact(env, n, d):
    rng = np.random.default_rng(n + d)
    u = rng.random((n, d))
    if n > 10:
        u[::7] = u[3]                                # heavy ties
        u[5] = 1e-300
    rows = rng.permutation(n)[: max(1, (2 * n) // 3)].astype(np.int64)
    mult = rng.integers(0, 5, size=rows.size).astype(np.int32)
    mult[0] = max(mult[0], 2)
    expanded = np.repeat(u[rows], mult, axis=0)
    srt = np.sort(expanded, axis=0)
    m_total = expanded.shape[0]
    du, dr, dm = dev_arr(env, u), dev_arr(env, rows), dev_arr(env, mult)
    out = torch.empty(2 * d, dtype=torch.float64, device=env.dev)
    for rank_lo in sorted({0, m_total // 2 - 1, m_total - 2, m_total - 1} & set(range(m_total))):
        same = rank_lo == m_total - 1
        env.k.g_select_pair(du, dr, d, rows.size, d, dm, rank_lo, same, out)
        want = np.stack([srt[rank_lo], srt[rank_lo if same else rank_lo + 1]], axis=1)
        np.testing.assert_array_equal(out.cpu().numpy().reshape(d, 2), want, err_msg=f"rank {rank_lo}")


@pytest.mark.parametrize("n,d,width", [(3, 1, 1.0), (5000, 10, 1.0), (200000, 10, 0.01), (50000, 4, 1e-6), (4000, 2, 0.0)])
def test_unit_median_pair_exact(env, n, d, width):
    """np.median's middle pair of the count-expanded rows (student.py:62), incl. concentrated and
    fully degenerate columns (every value in one bucket)."""
    rng = np.random.default_rng(n * 7 + d)
    u = 0.4 + width * rng.random((n, d)) * 0.5
    u[0] = 0.0
    u[-1] = 1.0
    rows = rng.permutation(n)[: max(2, (3 * n) // 4)].astype(np.int64)
    mult = rng.integers(0, 6, size=rows.size).astype(np.int32)
    mult[:2] = 1
    if mult.sum() % 2:
        mult[0] += 1
    m_total = int(mult.sum())
    srt = np.sort(np.repeat(u[rows], mult, axis=0), axis=0)
    du, dr, dm = dev_arr(env, u), dev_arr(env, rows), dev_arr(env, mult)
    out = torch.empty(2 * d, dtype=torch.float64, device=env.dev)
    ovf = torch.zeros(1, dtype=torch.int32, device=env.dev)
    ws = torch.zeros(env.lib.tb_unit_median_workspace_bytes(d), dtype=torch.uint8, device=env.dev)
    r = m_total // 2 - 1
    env._lib.check(env.lib.tb_unit_median_pair(env.ptr(du), env.ptr(dr), env.ptr(dm), rows.size, d, r, env.ptr(ws),
                                               env.ptr(out), env.ptr(ovf), env.sp()))
    if int(ovf.item()):
        assert rows.size * 1.0 > 65536          # only legitimate when a bucket really is that crowded
        env.k.g_select_pair(du, dr, d, rows.size, d, dm, r, False, out)
    np.testing.assert_array_equal(out.cpu().numpy().reshape(d, 2), np.stack([srt[r], srt[r + 1]], axis=1))
    assert np.array_equal(np.median(np.repeat(u[rows], mult, axis=0), axis=0), out.cpu().numpy().reshape(d, 2).mean(1))


# ---- hierarchical Gaussian-mixture clustering (cluster.py) against the oracle --------------------------
def _cluster_cases():
    from oracle.gen_golden import cluster_cases

    return cluster_cases()


@pytest.mark.parametrize("name", ["cluster_blobs2d", "cluster_blobs5d", "cluster_single10d", "cluster_raw3d_cap",
                                  "cluster_thin2d"])
def test_hierarchical_mixture_matches_oracle(env, name):
    from oracle import cluster_oracle as co
    from oracle import ps_oracle as po
    from tempest_b200.cluster import HierarchicalGaussianMixture, _Cluster

    x, w, kw = _cluster_cases()[name]
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    ref = co.fit_hierarchy(x, w, po.LegacyStream(0), **kw)
    k = env.k
    xd, wd = dev_arr(env, x), dev_arr(env, w)
    h = HierarchicalGaussianMixture(k, **kw).fit(xd, wd)
    assert h.n_clusters_ == ref.n_clusters == int(g["n_clusters"])
    np.testing.assert_array_equal(h.labels_.cpu().numpy(), g["labels"])
    np.testing.assert_allclose(np.array(h.cluster_centers_), g["centres"], rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(np.array(h.cluster_covariances_), g["covs"], rtol=1e-7, atol=1e-15)
    np.testing.assert_allclose(h.cluster_weights_, g["weights"], rtol=1e-12)
    np.testing.assert_array_equal(h.predict(dev_arr(env, g["y"])).cpu().numpy(), g["predict_y"])
    np.testing.assert_array_equal(h.predict(xd).cpu().numpy(), g["predict_x"])
    # the same fit through a row list (the trimmed-set form the sampler uses)
    perm = np.random.RandomState(1).permutation(len(x))
    big = np.zeros((len(x) + 7, x.shape[1]))
    big[perm] = x
    rows = dev_arr(env, perm.astype(np.int64), torch.int64)
    h2 = HierarchicalGaussianMixture(k, **kw).fit(dev_arr(env, big), wd, rows=rows)
    np.testing.assert_array_equal(h2.labels_.cpu().numpy(), g["labels"])
    # single mixtures: EM iteration counts, BIC and parameters (cluster.py:56-133, 310-340)
    hh = HierarchicalGaussianMixture(k, normalize=False)
    xn = xd.contiguous()
    cl = _Cluster(None, len(x))
    hh._weights_of(cl, wd)
    for comps in (1, 2):
        fit = hh._fit_mixture(xn, cl, comps, x.shape[1])
        assert fit.n_iter == int(g[f"gmm{comps}_n_iter"])
        assert fit.bic() == pytest.approx(float(g[f"gmm{comps}_bic"]), rel=1e-9)
        wts, means, covs = fit.host_params()
        np.testing.assert_allclose(wts, g[f"gmm{comps}_weights"], rtol=1e-8)
        np.testing.assert_allclose(means, g[f"gmm{comps}_means"], rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(covs, g[f"gmm{comps}_covs"], rtol=1e-6, atol=1e-14)


def test_split_by_label_keeps_order(env):
    rs = np.random.RandomState(4)
    for n in (1, 5, 2048, 2049, 100_003):
        labels = (rs.rand(n) < 0.37).astype(np.int32)
        members = np.sort(rs.choice(10 * n, n, replace=False)).astype(np.int64)
        ld, md = dev_arr(env, labels, torch.int32), dev_arr(env, members, torch.int64)
        zero = torch.empty(n, dtype=torch.int64, device=env.dev)
        one = torch.empty(n, dtype=torch.int64, device=env.dev)
        cnt = torch.zeros(1, dtype=torch.int64, device=env.dev)
        ws = torch.zeros(env.lib.tb_split_workspace_bytes(n) + 64, dtype=torch.uint8, device=env.dev)
        assert env.lib.tb_split_by_label(env.ptr(ld), env.ptr(md), n, env.ptr(ws), env.ptr(zero), env.ptr(one),
                                         env.ptr(cnt), env.sp()) == 0
        c1 = int(cnt.item())
        np.testing.assert_array_equal(one[:c1].cpu().numpy(), members[labels != 0])
        np.testing.assert_array_equal(zero[:n - c1].cpu().numpy(), members[labels == 0])


def test_guided_search_equals_plain_search(env):
    """tb_search_right_guided must return exactly the indices of the full binary search, including draws
    that sit on guide grid points, on cdf values, and in flat (zero-weight) stretches."""
    rng = np.random.default_rng(12)
    n, m = 200_000, 300_000
    p = np.exp(-0.5 * rng.chisquare(4, n) * 9.0)
    p[rng.integers(0, n, n // 10)] = 0.0
    p[1000:30000] = 0.0
    p /= p.sum()
    dp = dev_arr(env, p)
    cdf = env.k.cdf(dp, n, "g_cdf")
    c = cdf.cpu().numpy()
    u = rng.random(m)
    u[:4096] = np.arange(4096) / 4096.0                       # guide grid points of every table size
    u[4096:8192] = np.minimum(c[rng.integers(0, n, 4096)] / c[-1], np.nextafter(1.0, 0.0))   # exactly on cdf values
    u[8192] = np.nextafter(1.0, 0.0)
    du = dev_arr(env, u)
    plain = torch.empty(m, dtype=torch.int64, device=env.dev)
    assert env.lib.tb_search_right(env.ptr(cdf), n, env.ptr(du), m, env.ptr(plain), env.sp()) == 0
    ref = (c / c[-1]).searchsorted(u, side="right")
    np.testing.assert_array_equal(plain.cpu().numpy(), ref)
    for bits in (4, 10, 15, 20):
        guide = torch.zeros(env.lib.tb_search_guide_bytes(bits), dtype=torch.uint8, device=env.dev)
        out = torch.empty(m, dtype=torch.int64, device=env.dev)
        assert env.lib.tb_search_right_guided(env.ptr(cdf), n, env.ptr(du), m, env.ptr(guide), bits, env.ptr(out),
                                              env.sp()) == 0
        np.testing.assert_array_equal(out.cpu().numpy(), ref, err_msg=f"bits={bits}")


def test_scale_inplace_dev_divides_by_device_scalar(env):
    rng = np.random.default_rng(21)
    w = rng.random(100003)
    dw = dev_arr(env, w)
    den = dev_arr(env, np.array([0.0, 3.7, 0.0]))
    env._lib.check(env.lib.tb_scale_inplace_dev(env.ptr(dw), len(w), env.ptr(den[1:]), env.sp()))
    assert np.array_equal(dw.cpu().numpy(), w / 3.7)


def test_bucket_merge_concatenates_rank_lists_in_rank_order(env):
    """Sharded order statistics: the all-gathered candidate lists of the ranks are merged on the device."""
    G, d, capx, cap = 3, 2, 8, 65536
    rng = np.random.default_rng(22)
    counts = np.array([[3, 0], [8, 5], [1, 2]], dtype=np.int32)             # [G][d]
    gsel = np.zeros((G, d, 2), dtype=np.int32)
    gsel[:, :, 0] = counts
    gval = rng.random((G, d, capx))
    gmul = rng.integers(1, 9, size=(G, d, capx)).astype(np.int32)
    off = np.zeros(4, dtype=np.int64)
    env._lib.check(env.lib.tb_bucket_offsets(d, off.ctypes.data))
    o0, o1, o2, o3 = [int(v) for v in off]
    ws = torch.zeros(int(env.lib.tb_unit_median_workspace_bytes(d)), dtype=torch.uint8, device=env.dev)
    a, b, c = dev_arr(env, gsel), dev_arr(env, gval), dev_arr(env, gmul)
    env._lib.check(env.lib.tb_bucket_merge(env.ptr(a), env.ptr(b), env.ptr(c), G, d, capx, env.ptr(ws), env.sp()))
    sel = ws[o1: o1 + 24 * d].view(torch.int32).reshape(d, 6).cpu().numpy()
    cval = ws[o2: o2 + 8 * d * cap].view(torch.float64).reshape(d, cap).cpu().numpy()
    cmul = ws[o3: o3 + 4 * d * cap].view(torch.int32).reshape(d, cap).cpu().numpy()
    for col in range(d):
        want_v = np.concatenate([gval[r, col, : counts[r, col]] for r in range(G)])
        want_m = np.concatenate([gmul[r, col, : counts[r, col]] for r in range(G)])
        assert sel[col, 4] == len(want_v) and sel[col, 5] == 0
        np.testing.assert_array_equal(cval[col, : len(want_v)], want_v)
        np.testing.assert_array_equal(cmul[col, : len(want_m)], want_m)
    gsel[1, 0, 0] = capx + 1                                                 # a rank with more candidates than slots
    a = dev_arr(env, gsel)
    env._lib.check(env.lib.tb_bucket_merge(env.ptr(a), env.ptr(b), env.ptr(c), G, d, capx, env.ptr(ws), env.sp()))
    sel = ws[o1: o1 + 24 * d].view(torch.int32).reshape(d, 6).cpu().numpy()
    assert sel[0, 5] == 1 and sel[1, 5] == 0


def test_xgpu_bench_grid_reduction_checksum(env):
    """The in-kernel synchronisation point of the persistent kernels (grid_xreduce) on one GPU: every CTA must see the
    sum over all CTAs at every one of `count` back-to-back synchronisation points."""
    grid, count = 2 * int(env.lib.tb_sm_count()), 200
    ws = torch.zeros(int(env.lib.tb_xgpu_bench_workspace_bytes(grid)) + 256, dtype=torch.uint8, device=env.dev)
    out = torch.zeros(2, dtype=torch.float64, device=env.dev)
    env._lib.check(env.lib.tb_xgpu_bench(None, grid, count, env.ptr(ws), env.ptr(out), env.sp()))
    h = out.cpu().numpy()
    assert h[1] == float(sum(k & 7 for k in range(count)) * grid)
    assert 0.0 < h[0] < 1e6                                                  # ns per synchronisation point
