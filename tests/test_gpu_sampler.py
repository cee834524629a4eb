"""End-to-end parity of the device sampler with the CPU oracle (and, through the committed
golden fixtures, with the unmodified reference) on shared random-variate tapes.

The oracle is run on numpy's legacy MT19937 stream (the reference's own stream), records every
variate it consumes per PS iteration, and the CUDA path replays those tapes.  Discrete outputs
(beta sequence, trim set, training draws, resampling indices, MCMC step counts, accept
decisions) must be identical; continuous outputs must agree to 1e-10 relative (north star).
"""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from conftest import GOLDEN  # noqa: E402

RTOL = 1e-10


# N = 21: after two warm-up generations the reference's ESS of 42 exactly equal weights evaluates to
# 42.000000000000007 > target, so it leaves the warm-up and bisects beta in a region (beta ~ 1e-10)
# where every `ESS >= target` comparison is decided by the last-bit rounding of numpy's exp / pairwise
# sums (the true ESS deficit there is ~1e-16).  No independent implementation can reproduce those
# bits (parity hazards); the device path must take the same branch out of the warm-up and
# agree on everything downstream to the accuracy that the ~1e-10 absolute beta difference allows.
ROUNDING_DECIDED = {"mix2_n21_tpcn_mult"}


def check_rounding_decided_case(o, s):
    st = s.state
    b, bo = st.get_history("beta"), np.array(o.hist["beta"])
    assert np.array_equal(b == 0.0, bo == 0.0)                       # same warm-up exit
    np.testing.assert_allclose(b, bo, rtol=1e-6, atol=1e-9)
    np.testing.assert_array_equal(st.get_history("steps"), np.array(o.hist["steps"]))
    np.testing.assert_allclose(st.get_history("logz"), np.array(o.hist["logz"]), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(st.get_history("u"), np.array(o.hist["u"]), rtol=1e-6, atol=1e-9)
    assert s.evidence()[0] == pytest.approx(o.evidence()[0], rel=1e-6)


def run_pair(name, max_iterations=None):
    """Oracle on LegacyStream(seed) with recording, then the device sampler on its tapes."""
    import tempest_b200 as tp
    from oracle import ps_oracle as po
    from oracle.gen_golden import cases
    from tempest_b200.rng import TapeSource

    prior, like, kw, n_total, seed = cases()[name]
    o = po.OraclePS(prior, like, stream=po.LegacyStream(seed), record=True, **kw)
    o.run(n_total, max_iterations=max_iterations)
    s = tp.Sampler(prior, like, vectorize=True, **kw)
    core = s._core
    core.rng = TapeSource(o.tapes, core.device)
    core._initialize_fresh()
    core.n_total = int(n_total)
    traces = []
    k = 0
    while core._not_termination():
        core.execute_iteration()
        tr = {key: (v.cpu().numpy() if hasattr(v, "cpu") else v) for key, v in core.trace.items()}
        for key, v in tr.items():
            if isinstance(v, list):
                tr[key] = [e.cpu().numpy() if hasattr(e, "cpu") else e for e in v]
        if core.clusterer is not None and core.clusterer.n_clusters_:
            tr["n_clusters"] = core.clusterer.n_clusters_
            tr["cluster_centres"] = np.array(core.clusterer.cluster_centers_)
            tr["cluster_covs"] = np.array(core.clusterer.cluster_covariances_)
            tr["cluster_weights"] = np.array(core.clusterer.cluster_weights_)
        tr["probe_log"] = list(core.reweighter.probe_log)
        tr["mode_mean"] = core.last_mode_stats.means.cpu().numpy()
        tr["mode_cov"] = core.last_mode_stats.covariances.cpu().numpy()
        traces.append(tr)
        k += 1
        if max_iterations is not None and k >= max_iterations:
            break
    core.state.set_current("logz", float(core._last_posterior_probe[4]))
    return o, s, traces


@pytest.mark.parametrize("name", ["rosen10_n64_tpcn_mult", "gauss4_n32_rwm_syst_bc", "mix2_n21_tpcn_mult",
                                  "shell4_n48_dynamic"])
def test_full_run_matches_oracle_on_tapes(name):
    o, s, traces = run_pair(name)
    st = s.state
    T = len(o.hist["beta"])
    assert st.get_history_length() == T
    if name in ROUNDING_DECIDED:
        return check_rounding_decided_case(o, s)
    # discrete: beta sequence, step counts, call counts
    np.testing.assert_array_equal(st.get_history("beta"), np.array(o.hist["beta"]))
    np.testing.assert_array_equal(st.get_history("steps"), np.array(o.hist["steps"]))
    np.testing.assert_array_equal(st.get_history("calls"), np.array(o.hist["calls"]))
    np.testing.assert_array_equal(st.get_history("iter"), np.array(o.hist["iter"]))
    for t, (tr, otr) in enumerate(zip(traces, o.traces)):
        for key in ("trim_idx", "train_draw_idx", "resample_idx"):
            if key in otr:
                np.testing.assert_array_equal(tr[key], otr[key], err_msg=f"{key} @ iteration {t}")
        # probe sequence (the oracle re-probes the final beta when the bracket collapses)
        mine = np.array(tr["probe_log"]).reshape(-1, 2)
        ref = np.array(otr["probes"]).reshape(-1, 2)
        if len(ref) == len(mine) + 1:
            ref = ref[:-1]
        np.testing.assert_array_equal(mine[:, 0], ref[:, 0], err_msg=f"probe betas @ iteration {t}")
        np.testing.assert_allclose(mine[:, 1], ref[:, 1], rtol=RTOL)
        if "mode_mean" in otr:
            # medians are exact order statistics of walker coordinates that already differ in the last bits
            np.testing.assert_allclose(tr["mode_mean"], otr["mode_mean"], rtol=1e-12)
            np.testing.assert_allclose(tr["mode_cov"], otr["mode_cov"], rtol=1e-9, atol=1e-14)
    # continuous: every generation of particles, evidence, ESS, cv, acceptance
    np.testing.assert_allclose(st.get_history("logz"), np.array(o.hist["logz"]), rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(st.get_history("ess"), np.array(o.hist["ess"]), rtol=RTOL)
    np.testing.assert_allclose(st.get_history("cv"), np.array(o.hist["cv"]), rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(st.get_history("acceptance"), np.array(o.hist["acceptance"]), rtol=RTOL)
    np.testing.assert_allclose(st.get_history("efficiency"), np.array(o.hist["efficiency"]), rtol=RTOL)
    np.testing.assert_allclose(st.get_history("u"), np.array(o.hist["u"]), rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(st.get_history("x"), np.array(o.hist["x"]), rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(st.get_history("logl"), np.array(o.hist["logl"]), rtol=1e-9, atol=1e-9)
    # accept/reject decisions: a flipped decision would leave a walker at a different point
    du = np.abs(st.get_history("u") - np.array(o.hist["u"])).max()
    assert du < 1e-12, f"walker positions differ by {du}: an accept/reject decision flipped"
    assert s.evidence()[0] == pytest.approx(o.evidence()[0], rel=RTOL)
    # posterior() against the oracle and against the reference's own output (golden fixture)
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    x, w, l, logw = s.posterior(return_logw=True)
    assert x.shape == g["post_x"].shape
    np.testing.assert_allclose(x, g["post_x"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(w, g["post_w"], rtol=1e-9)
    np.testing.assert_allclose(l, g["post_logl"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(logw, g["post_logw"], rtol=1e-9, atol=1e-9)
    assert s.evidence()[0] == pytest.approx(float(g["final_logz"]), rel=RTOL)
    np.testing.assert_array_equal(st.get_history("beta"), g["h_beta"])
    np.testing.assert_array_equal(st.get_history("steps"), g["h_steps"])


@pytest.mark.parametrize("name", ["mix2_n64_clustered", "mix3_n96_clustered_cap"])
def test_clustered_run_matches_oracle_on_tapes(name):
    """clustering=True: hierarchical mixture fits on the device (csrc/tb_cluster.cu) must find the same
    clusters, labels and walker assignments as the oracle (pinned bit-exact to the reference), and
    the multi-mode MCMC must then take the same steps."""
    o, s, traces = run_pair(name)
    st = s.state
    assert st.get_history_length() == len(o.hist["beta"])
    np.testing.assert_array_equal(st.get_history("beta"), np.array(o.hist["beta"]))
    np.testing.assert_array_equal(st.get_history("steps"), np.array(o.hist["steps"]))
    np.testing.assert_array_equal(st.get_history("calls"), np.array(o.hist["calls"]))
    seen_multi = False
    for t, (tr, otr) in enumerate(zip(traces, o.traces)):
        if "train_labels" not in otr:
            continue
        assert tr["n_clusters"] == otr["n_clusters"], f"cluster count @ iteration {t}"
        seen_multi |= otr["n_clusters"] > 1
        np.testing.assert_array_equal(tr["trim_idx"], otr["trim_idx"], err_msg=f"trim @ {t}")
        np.testing.assert_array_equal(tr["train_labels"], otr["train_labels"], err_msg=f"labels @ {t}")
        np.testing.assert_allclose(tr["cluster_centres"], otr["cluster_centres"], rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(tr["cluster_covs"], otr["cluster_covs"], rtol=1e-7, atol=1e-14)
        np.testing.assert_allclose(tr["cluster_weights"], otr["cluster_weights"], rtol=1e-10)
        assert len(tr["train_draw_idx"]) == len(otr["train_draw_idx"])
        for a, b in zip(tr["train_draw_idx"], otr["train_draw_idx"]):
            np.testing.assert_array_equal(a, b, err_msg=f"mode draws @ {t}")
        np.testing.assert_array_equal(tr["resample_idx"], otr["resample_idx"], err_msg=f"resample @ {t}")
        np.testing.assert_array_equal(tr["assignments"], otr["assignments"], err_msg=f"assignments @ {t}")
        np.testing.assert_allclose(tr["mode_mean"], otr["mode_mean"], rtol=1e-12)
        np.testing.assert_allclose(tr["mode_cov"], otr["mode_cov"], rtol=1e-9, atol=1e-14)
    assert seen_multi, "the case never produced more than one cluster: it does not exercise the hierarchy"
    np.testing.assert_allclose(st.get_history("logz"), np.array(o.hist["logz"]), rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(st.get_history("acceptance"), np.array(o.hist["acceptance"]), rtol=RTOL)
    np.testing.assert_allclose(st.get_history("efficiency"), np.array(o.hist["efficiency"]), rtol=RTOL)
    du = np.abs(st.get_history("u") - np.array(o.hist["u"])).max()
    assert du < 1e-12, f"walker positions differ by {du}: an accept/reject decision flipped"
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    np.testing.assert_array_equal(st.get_history("beta"), g["h_beta"])
    np.testing.assert_array_equal(st.get_history("steps"), g["h_steps"])
    assert s.evidence()[0] == pytest.approx(float(g["final_logz"]), rel=RTOL)


def test_generic_step_kernel_matches_oracle_on_tapes():
    """The runtime-d kernel (used for n_dim without a compile-time instantiation) on the same tapes."""
    from tempest_b200 import _lib

    lib = _lib.load()
    lib.tb_set_mcmc_generic(1)
    try:
        o, s, _ = run_pair("rosen10_n64_tpcn_mult", max_iterations=12)
    finally:
        lib.tb_set_mcmc_generic(0)
    np.testing.assert_array_equal(s.state.get_history("beta"), np.array(o.hist["beta"]))
    np.testing.assert_array_equal(s.state.get_history("steps"), np.array(o.hist["steps"]))
    np.testing.assert_allclose(s.state.get_history("u"), np.array(o.hist["u"]), rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(s.state.get_history("logl"), np.array(o.hist["logl"]), rtol=1e-9, atol=1e-9)


def test_multimode_instantiation_on_single_mode_runs_matches_oracle_on_tapes():
    """Single-mode runs normally take the constant-memory instantiation <D, ., TAPE, KONE=1, LIKE> of the fused step
    kernel -- the one bench.py times, run under the tapes by every test above.  This routes the same run through the
    multi-mode instantiation (shared-memory operands, runtime likelihood switch): both must take the oracle's decisions."""
    from tempest_b200 import _lib

    lib = _lib.load()
    lib.tb_set_mcmc_kone(0)
    try:
        o, s, _ = run_pair("rosen10_n64_tpcn_mult", max_iterations=12)
    finally:
        lib.tb_set_mcmc_kone(1)
    np.testing.assert_array_equal(s.state.get_history("beta"), np.array(o.hist["beta"]))
    np.testing.assert_array_equal(s.state.get_history("steps"), np.array(o.hist["steps"]))
    np.testing.assert_allclose(s.state.get_history("u"), np.array(o.hist["u"]), rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(s.state.get_history("logl"), np.array(o.hist["logl"]), rtol=1e-9, atol=1e-9)
    assert np.abs(s.state.get_history("u") - np.array(o.hist["u"])).max() < 1e-12


@pytest.mark.parametrize("like", ["rosenbrock", "gaussian", "mixture", "shells"])
def test_every_compile_time_likelihood_variant_matches_oracle_on_tapes(like):
    """One short tape run per registry likelihood at n_dim = 10: each selects its own <10, tpCN, TAPE, KONE, LIKE>
    instantiation (the production kernels differ from these only in the source of the three variates)."""
    import tempest_b200 as tp
    from oracle import ps_oracle as po
    from tempest_b200.rng import TapeSource

    d, n, iters = 10, 96, 9
    prior = tp.UniformPrior(-6.0, 6.0, d)
    if like == "rosenbrock":
        L = tp.Rosenbrock(d)
    elif like == "gaussian":
        L = tp.GaussianLikelihood.ar1(d, 0.5)
    elif like == "mixture":
        rng = np.random.default_rng(3)
        L = tp.IsotropicMixture(rng.uniform(-3, 3, (3, d)), [0.6, 0.9, 0.7], [0.2, 0.5, 0.3])
    else:
        L = tp.TwinShells(d)
    o = po.OraclePS(prior, L, d, n_particles=n, stream=po.LegacyStream(31), record=True)
    o.run(1 << 30, max_iterations=iters)
    s = tp.Sampler(prior, L, d, n_particles=n, vectorize=True, clustering=False)
    core = s._core
    core.rng = TapeSource(o.tapes, core.device)
    core._initialize_fresh()
    for _ in range(iters):
        core.execute_iteration(export=False)
    st = s.state
    np.testing.assert_array_equal(st.get_history("beta"), np.array(o.hist["beta"]))
    np.testing.assert_array_equal(st.get_history("steps"), np.array(o.hist["steps"]))
    np.testing.assert_allclose(st.get_history("logz"), np.array(o.hist["logz"]), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(st.get_history("logl"), np.array(o.hist["logl"]), rtol=1e-9, atol=1e-9)
    assert np.abs(st.get_history("u") - np.array(o.hist["u"])).max() < 1e-12      # no accept / reject flip


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharded_run_matches_single_gpu():
    """Sharded over 2 ranks (NCCL) the run must reproduce the 1-GPU beta sequence, step counts and logZ."""
    import subprocess
    import sys

    from conftest import ROOT

    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533",
                          os.path.join(ROOT, "tools", "dist_check.py"), "4096"],
                         capture_output=True, text=True, timeout=600)
    assert "DIST OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_host_driven_search_equals_device_search():
    import tempest_b200 as tp

    outs = []
    for device_search in (True, False):
        s = tp.Sampler(tp.UniformPrior(-10, 10, 10), tp.Rosenbrock(10), 10, n_particles=512, vectorize=True,
                       clustering=False, random_state=5)
        s._core.reweighter.device_search = device_search
        s._core._initialize_fresh()
        for _ in range(8):
            s.sample()
        outs.append((s.state.get_history("beta"), s.state.get_history("logz"), s.state.get_history("logl")))
    for a, b in zip(*outs):
        np.testing.assert_array_equal(a, b)


def test_philox_run_is_reproducible_and_recovers_gaussian_evidence():
    # reference tests/test_end_to_end.py: 10-D diagonal Gaussian, U(-10,10) prior, logZ = -29.96 +- 0.5
    import tempest_b200 as tp

    mean = np.array([2.0, -1.5, 0.5, 3.2, -2.8, 1.1, -0.7, 2.5, -1.2, 0.9])
    var = np.array([1.0, 0.8, 1.2, 0.9, 1.1, 0.7, 1.3, 0.85, 1.15, 0.95])

    def run():
        s = tp.Sampler(tp.UniformPrior(-10, 10, 10), tp.GaussianLikelihood(mean, var), 10, n_particles=1024,
                       vectorize=True, clustering=False, random_state=42, n_steps=1)
        s.run(n_total=4096, progress=False)
        return s

    s = run()
    logz, err = s.evidence()
    assert err is None and abs(logz - (-29.96)) < 0.5
    x, w, l = s.posterior()
    assert x.shape[1] == 10 and x.shape[0] == w.shape[0] == l.shape[0]
    assert w.sum() == pytest.approx(1.0, abs=1e-12)
    pm = np.average(x, weights=w, axis=0)
    np.testing.assert_allclose(pm, mean, atol=0.25)
    pc = np.cov(x, rowvar=False, aweights=w)
    np.testing.assert_allclose(np.diag(pc), var, atol=0.5)
    assert s.beta > 0.99 and s.state.get_current("acceptance") > 0.1
    s2 = run()
    np.testing.assert_array_equal(s2.state.get_history("beta"), s.state.get_history("beta"))
    np.testing.assert_array_equal(s2.state.get_history("logl", flat=True), s.state.get_history("logl", flat=True))
    assert s2.evidence()[0] == logz
    # resampled posterior has uniform weights (core.py:222-231)
    xr, wr, lr = s.posterior(resample=True)
    assert np.all(wr == 1.0 / len(wr))


def test_sample_contract_and_state_surface():
    # reference tests/test_sample_method.py:80-146, tests/test_state_manager.py:204-281
    import tempest_b200 as tp
    from tempest_b200.ensemble import CURRENT_STATE_KEYS

    s = tp.Sampler(tp.UniformPrior(-10, 10, 4), tp.Rosenbrock(4), 4, n_particles=32, vectorize=True,
                   clustering=False, random_state=1)
    logw, logz = s.state.compute_logw_and_logz(1.0)
    assert logw.size == 0 and logz == -np.inf
    out = s.sample()
    assert set(out) == set(CURRENT_STATE_KEYS)
    assert out["u"].shape == (32, 4) and out["x"].shape == (32, 4) and out["logl"].shape == (32,)
    assert out["beta"] == 0.0 and out["ess"] == 64.0 and out["iter"] == 1 and out["calls"] == 32
    out["u"][:] = -1.0                                       # copies, not views
    assert s.state.get_current("u").min() >= 0.0
    for _ in range(4):
        s.sample()
    assert s.state.get_history_length() == 5
    assert s.state.get_history("u").shape == (5, 32, 4)
    assert s.state.get_history("logl", flat=True).shape == (160,)
    logw, logz = s.state.compute_logw_and_logz(1.0)
    assert np.exp(logw).sum() == pytest.approx(1.0, abs=1e-9) and np.isfinite(logz)
    assert s.n_dim == 4 and s.n_particles == 32 and s.resample == "mult" and s.clustering is False
    with pytest.raises(NotImplementedError):      # blobs are not carried by the device ensemble
        tp.Sampler(tp.UniformPrior(-1, 1, 2), lambda x: -np.sum(x * x), 2, clustering=False, blobs_dtype="float")
    with pytest.raises(ValueError, match="Invalid resample"):
        tp.Sampler(tp.UniformPrior(-1, 1, 2), tp.Rosenbrock(2), 2, vectorize=True, clustering=False, resample="x")


@pytest.mark.parametrize("case", ["gauss50_pcn", "shells100"])
def test_high_dimensional_configs_match_oracle_on_tapes(case):
    """BASELINE configs[2] / [4] shapes (50-D correlated Gaussian with pCN; 100-D twin shells) at small N:
    exercises the runtime-d step kernel, the smem-tile moment kernels and d x d Cholesky at d = 50 / 100."""
    import tempest_b200 as tp
    from oracle import ps_oracle as po
    from tempest_b200.rng import TapeSource

    if case == "gauss50_pcn":
        d, n, iters = 50, 128, 7
        prior, like = tp.UniformPrior(-10.0, 10.0, d), tp.GaussianLikelihood.ar1(d, 0.5)
    else:
        d, n, iters = 100, 256, 6
        prior, like = tp.UniformPrior(-6.0, 6.0, d), tp.TwinShells(d)
    o = po.OraclePS(prior, like, d, n_particles=n, stream=po.LegacyStream(77), record=True)
    o.run(1 << 30, max_iterations=iters)
    s = tp.Sampler(prior, like, d, n_particles=n, vectorize=True, clustering=False)
    core = s._core
    core.rng = TapeSource(o.tapes, core.device)
    core._initialize_fresh()
    for _ in range(iters):
        core.execute_iteration(export=False)
    st = s.state
    np.testing.assert_array_equal(st.get_history("beta"), np.array(o.hist["beta"]))
    np.testing.assert_array_equal(st.get_history("steps"), np.array(o.hist["steps"]))
    np.testing.assert_allclose(st.get_history("logz"), np.array(o.hist["logz"]), rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(st.get_history("cv"), np.array(o.hist["cv"]), rtol=1e-7, atol=1e-10)
    np.testing.assert_allclose(st.get_history("u"), np.array(o.hist["u"]), rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(st.get_history("logl"), np.array(o.hist["logl"]), rtol=1e-8, atol=1e-8)
    assert np.abs(st.get_history("u") - np.array(o.hist["u"])).max() < 1e-10      # no accept/reject flip


def test_readme_default_config_with_clustering_runs_to_the_posterior():
    """Reference README.md:44-71 verbatim (config C1): n_particles defaults to 2*n_dim = 20 and
    clustering is on.  The reference's own run gives logZ in [-29.9, -29.5] for this likelihood; with
    20 particles the estimate is noisy, so the check is a loose statistical one."""
    import tempest_b200 as tp

    n_dim = 10
    s = tp.Sampler(prior_transform=tp.UniformPrior(-10.0, 10.0, n_dim), log_likelihood=tp.Rosenbrock(n_dim),
                   n_dim=n_dim, vectorize=True, random_state=0)
    assert s.n_particles == 20 and s.clustering is True
    s.run(progress=False)
    samples, weights, logl = s.posterior()
    logz, logz_err = s.evidence()
    assert samples.shape[1] == n_dim and samples.shape[0] == weights.shape[0] == logl.shape[0]
    assert weights.sum() == pytest.approx(1.0, abs=1e-9)
    assert -33.0 < logz < -27.0
    assert s.state.get_history_length() > 50
    assert s._core.clusterer.n_fits > 0


def test_clustered_mixture_recovers_the_analytic_evidence():
    """Config C2 shape at N = 4096: four separated 2-D modes, clustering on, logZ = -log 400."""
    import tempest_b200 as tp

    s = tp.Sampler(tp.UniformPrior(-10.0, 10.0, 2), tp.IsotropicMixture.four_corners(2), 2, n_particles=4096,
                   vectorize=True, clustering=True, random_state=5)
    s.run(progress=False)
    assert s.evidence()[0] == pytest.approx(-np.log(400.0), abs=0.08)
    x, w, _ = s.posterior()
    # each of the four modes holds a quarter of the posterior mass
    for sx in (-1, 1):
        for sy in (-1, 1):
            mass = w[(np.sign(x[:, 0]) == sx) & (np.sign(x[:, 1]) == sy)].sum()
            assert mass == pytest.approx(0.25, abs=0.05)


# ---- arbitrary user callables (core.py:317-358): split propose / accept step --------------------------
@pytest.mark.parametrize("name", ["rosen10_n64_tpcn_mult", "gauss4_n32_rwm_syst_bc", "mix2_n64_clustered"])
def test_plain_python_callables_match_oracle_on_tapes(name):
    """The same run as the fused path, but the prior and the likelihood are opaque numpy callables
    (closures around the registry objects, so the numbers are comparable): every decision must match
    the oracle exactly as it does for the in-kernel likelihoods."""
    import tempest_b200 as tp
    from oracle import ps_oracle as po
    from oracle.gen_golden import cases
    from tempest_b200.rng import TapeSource

    prior, like, kw, n_total, seed = cases()[name]
    o = po.OraclePS(prior, like, stream=po.LegacyStream(seed), record=True, **kw)
    o.run(n_total)
    calls = {"prior_rows": 0, "like": 0}

    def my_prior(u):
        calls["prior_rows"] += 1
        return prior(u)

    def my_like(x, shift, scale=1.0):
        calls["like"] += 1
        assert x.ndim == 2
        return (like(x) + shift) * scale

    s = tp.Sampler(my_prior, my_like, vectorize=True, log_likelihood_args=[0.0],
                   log_likelihood_kwargs={"scale": 1.0}, **kw)
    core = s._core
    assert core.bridge.external and core.bridge.prior_batched
    core.rng = TapeSource(o.tapes, core.device)
    core._initialize_fresh()
    core.n_total = int(n_total)
    while core._not_termination():
        core.execute_iteration()
    st = s.state
    np.testing.assert_array_equal(st.get_history("beta"), np.array(o.hist["beta"]))
    np.testing.assert_array_equal(st.get_history("steps"), np.array(o.hist["steps"]))
    np.testing.assert_array_equal(st.get_history("calls"), np.array(o.hist["calls"]))
    np.testing.assert_allclose(st.get_history("logz"), np.array(o.hist["logz"]), rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(st.get_history("logl"), np.array(o.hist["logl"]), rtol=1e-9, atol=1e-9)
    du = np.abs(st.get_history("u") - np.array(o.hist["u"])).max()
    assert du < 1e-12
    np.testing.assert_allclose(st.get_history("x"), np.array(o.hist["x"]), rtol=1e-12, atol=1e-12)
    assert calls["like"] >= int(np.sum(o.hist["steps"]))


def test_rowwise_prior_and_unvectorised_likelihood():
    """A prior written for one sample (indexes coordinates, so a batched call would be wrong) and
    vectorize=False: the bridge must fall back to per-row calls like the reference (mcmc.py:157,
    core.py:323-326) and still recover the analytic evidence of a unit Gaussian in a [-8, 8]^3 box."""
    import tempest_b200 as tp

    d = 3

    def prior(u):
        x = np.empty(d)
        x[0] = 16.0 * u[0] - 8.0
        x[1] = 16.0 * u[1] - 8.0
        x[2] = 16.0 * u[2] - 8.0
        return x

    def like(x):
        assert x.shape == (d,)
        return float(-0.5 * np.sum(x * x))

    s = tp.Sampler(prior, like, d, n_particles=256, vectorize=False, clustering=False, random_state=3)
    assert s._core.bridge.external and not s._core.bridge.prior_batched
    s.run(n_total=1024, progress=False)
    exact = 0.5 * d * np.log(2 * np.pi) - d * np.log(16.0)
    assert s.evidence()[0] == pytest.approx(exact, abs=0.25)
    x, w, logl = s.posterior()
    assert x.shape[1] == d and abs(np.average(x[:, 0], weights=w)) < 0.3


def test_device_callables_stay_on_the_gpu():
    """Callables marked with tp.device_callable get CUDA fp64 tensors; nothing crosses PCIe per step."""
    import tempest_b200 as tp

    d = 4
    seen = {"cuda": True}

    @tp.device_callable
    def prior(u):
        seen["cuda"] &= u.is_cuda
        return 20.0 * u - 10.0

    @tp.device_callable
    def like(x):
        seen["cuda"] &= x.is_cuda
        return -0.5 * (x * x).sum(dim=1)

    s = tp.Sampler(prior, like, d, n_particles=4096, vectorize=True, clustering=False, random_state=9)
    s.run(progress=False)
    exact = 0.5 * d * np.log(2 * np.pi) - d * np.log(20.0)
    assert seen["cuda"]
    assert s.evidence()[0] == pytest.approx(exact, abs=0.05)
    # same seed, registry objects in the fused kernel: statistically the same answer
    s2 = tp.Sampler(tp.UniformPrior(-10.0, 10.0, d), tp.GaussianLikelihood(np.zeros(d), np.ones(d)), d,
                    n_particles=4096, vectorize=True, clustering=False, random_state=9)
    s2.run(progress=False)
    const = -0.5 * d * np.log(2 * np.pi)           # the registry Gaussian is normalised
    assert s2.evidence()[0] - const == pytest.approx(s.evidence()[0], abs=0.05)


def test_checkpoint_resume_reproduces_the_uninterrupted_run(tmp_path):
    """save_every / resume_state_path (core.py:110-160, 249-315): a run resumed from a mid-run state file
    must finish exactly like the uninterrupted run (Philox counters are keyed by the PS iteration)."""
    import tempest_b200 as tp

    def make():
        return tp.Sampler(tp.UniformPrior(-10.0, 10.0, 4), tp.Rosenbrock(4), 4, n_particles=512, vectorize=True,
                          clustering=False, random_state=17, output_dir=str(tmp_path), output_label="ck")

    a = make()
    a.run(n_total=1024, progress=False, save_every=4)
    T = a.state.get_history_length()
    assert T > 9
    assert (tmp_path / "ck_4.state").exists() and (tmp_path / "ck_8.state").exists()
    assert (tmp_path / "ck_final.state").exists()
    b = make()
    b.run(n_total=1024, progress=False, resume_state_path=tmp_path / "ck_8.state")
    assert b.state.get_history_length() == T
    for key in ("beta", "logz", "steps", "calls", "ess", "acceptance"):
        np.testing.assert_array_equal(b.state.get_history(key), a.state.get_history(key), err_msg=key)
    np.testing.assert_array_equal(b.state.get_history("u"), a.state.get_history("u"))
    np.testing.assert_array_equal(b.state.get_history("logl"), a.state.get_history("logl"))
    assert b.evidence()[0] == a.evidence()[0]
    # save_state / load_state round trip of a finished run, and results()
    c = make()
    c.load_state(tmp_path / "ck_final.state")
    xa, wa, la = a.posterior()
    xc, wc, lc = c.posterior()
    np.testing.assert_array_equal(xa, xc)
    np.testing.assert_array_equal(wa, wc)
    res = c.results()
    assert res["logw"].shape == (T * 512,) and res["u"].shape == (T, 512, 4)


def test_core_reset_reruns_on_the_same_buffers():
    """SamplerCore.reset() (used by bench.py between runs): history cleared, device buffers kept, and the
    rerun reproduces the first run bit for bit."""
    import tempest_b200 as tp

    s = tp.Sampler(tp.UniformPrior(-10.0, 10.0, 4), tp.Rosenbrock(4), 4, n_particles=2048, vectorize=True,
                   clustering=False, random_state=23)
    s.run(n_total=1024, progress=False)
    beta, logz, u_last = s.state.get_history("beta"), s.evidence()[0], s.state.get_history("u")[-1]
    ptr_before = s._core.ensemble.u.data_ptr()
    s._core.reset()
    assert s.state.get_history_length() == 0 and s._core.ensemble.n_total == 0
    s._core.n_total = 1024
    while s._core._not_termination():
        s._core.execute_iteration(export=False)
    s._core.state.set_current("logz", float(s._core._last_posterior_probe[4]))
    assert s._core.ensemble.u.data_ptr() == ptr_before
    np.testing.assert_array_equal(s.state.get_history("beta"), beta)
    np.testing.assert_array_equal(s.state.get_history("u")[-1], u_last)
    assert s.evidence()[0] == logz
