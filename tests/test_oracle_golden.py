"""Pins the CPU oracle (oracle/ps_oracle.py) to the reference.

(1) tests/golden/*.npz hold outputs of the UNMODIFIED reference run under np.random.seed(s)
    (oracle/gen_golden.py); the oracle replayed on the same legacy MT19937 stream must
    reproduce every generation bit-for-bit.
(2) The reference's own known-answer vectors for this path.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import ps_oracle as po
from oracle import cluster_oracle as co
from oracle.gen_golden import cases, cluster_cases

CASES = cases()
CLUSTER_CASES = cluster_cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_reference_run(name):
    prior, like, kw, n_total, seed = CASES[name]
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    o = po.OraclePS(prior, like, stream=po.LegacyStream(seed), **kw)
    o.run(n_total)
    assert len(o.hist["beta"]) == len(g["h_beta"])
    for k in ("beta", "logz", "ess", "cv", "steps", "acceptance", "efficiency", "calls", "iter"):
        np.testing.assert_array_equal(np.array(o.hist[k], dtype=float), g["h_" + k], err_msg=k)
    for k in ("u", "x", "logl"):
        np.testing.assert_array_equal(np.array(o.hist[k]), g[k], err_msg=k)
    assert o.evidence()[0] == float(g["final_logz"])
    x, w, l, logw = o.posterior(return_logw=True)
    np.testing.assert_array_equal(x, g["post_x"])
    np.testing.assert_array_equal(w, g["post_w"])
    np.testing.assert_array_equal(l, g["post_logl"])
    np.testing.assert_array_equal(logw, g["post_logw"])
    _, w2, _ = o.posterior(trim_importance_weights=False)
    np.testing.assert_array_equal(w2, g["post_w_untrimmed"])


@pytest.mark.parametrize("name", sorted(CLUSTER_CASES))
def test_cluster_oracle_reproduces_reference_fit(name):
    """cluster.py restated (oracle/cluster_oracle.py) against the reference's own fits: the EM
    iteration counts, labels, centres, covariances, BICs and predictions must be identical."""
    x, w, kw = CLUSTER_CASES[name]
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    np.testing.assert_array_equal(x, g["x"])
    np.testing.assert_array_equal(w, g["w"])
    h = co.fit_hierarchy(x, w, po.LegacyStream(0), **kw)
    assert h.n_clusters == int(g["n_clusters"])
    np.testing.assert_array_equal(h.labels, g["labels"])
    np.testing.assert_array_equal(np.array(h.centres), g["centres"])
    np.testing.assert_array_equal(np.array(h.covs), g["covs"])
    np.testing.assert_array_equal(h.weights, g["weights"])
    np.testing.assert_array_equal(h.predict(g["y"]), g["predict_y"])
    np.testing.assert_array_equal(h.predict(x), g["predict_x"])
    for k in (1, 2):
        m = co.fit_mixture(x, w, k, po.LegacyStream(0))
        assert m.n_iter == int(g[f"gmm{k}_n_iter"])
        np.testing.assert_array_equal(m.weights, g[f"gmm{k}_weights"])
        np.testing.assert_array_equal(m.means, g[f"gmm{k}_means"])
        np.testing.assert_array_equal(m.covs, g[f"gmm{k}_covs"])
        assert m.bic(x) == float(g[f"gmm{k}_bic"])
        assert m.lower_bound == float(g[f"gmm{k}_lower"])
        np.testing.assert_array_equal(m.predict(x), g[f"gmm{k}_labels"])


def test_mvn_logpdf_matches_scipy():
    from scipy.stats import multivariate_normal

    rs = np.random.RandomState(9)
    for d in (1, 2, 5, 12):
        a = rs.randn(d, d)
        cov = a @ a.T / d + 1e-3 * np.eye(d)
        mean = rs.randn(d)
        x = rs.randn(64, d)
        np.testing.assert_array_equal(co.mvn_logpdf(x, mean, cov), multivariate_normal.logpdf(x, mean=mean, cov=cov))
    with pytest.raises(np.linalg.LinAlgError):
        co.mvn_logpdf(np.zeros((2, 2)), np.zeros(2), np.array([[1.0, 1.0], [1.0, 1.0]]))


def test_boundary_known_answers():
    # reference tests/test_mcmc.py:14-94,187-217
    np.testing.assert_array_almost_equal(
        po.boundary_map(np.array([0.5, 1.5, -0.5, 2.3]), periodic=[0, 1, 2, 3]), [0.5, 0.5, 0.5, 0.3])
    np.testing.assert_array_almost_equal(
        po.boundary_map(np.array([0.5, 1.2, -0.3, 1.8]), reflective=[0, 1, 2, 3]), [0.5, 0.8, 0.3, 0.2])
    r = po.boundary_map(np.array([0.5, 1.2, -0.3, 2.1]), periodic=[0, 1], reflective=[2, 3])
    np.testing.assert_array_almost_equal(r, [0.5, 0.2, 0.3, 0.1])
    np.testing.assert_array_almost_equal(
        po.boundary_map(np.array([3.7, -2.3, 5.1]), periodic=[0, 1, 2]), [0.7, 0.7, 0.1])
    np.testing.assert_array_almost_equal(
        po.boundary_map(np.array([2.3, -1.7, 3.5]), reflective=[0, 1, 2]), [0.3, 0.3, 0.5])
    u2 = np.array([[0.5, 1.2, -0.3], [0.8, 2.1, 0.4], [1.5, -0.5, 0.6]])
    r2 = po.boundary_map(u2, periodic=[0], reflective=[1, 2])
    np.testing.assert_array_almost_equal(r2[:, 0], [0.5, 0.8, 0.5])
    np.testing.assert_array_almost_equal(r2[:, 1], [0.8, 0.1, 0.5])
    assert po.inside_unit_cube(np.array([0.0, 0.5, 1.0, 0.3, 0.7]))
    assert not po.inside_unit_cube(np.array([0.5, 1.5, 0.3, -0.2]))
    assert po.inside_unit_cube(np.array([0.5, 1.5, 0.3]), periodic=[1])
    np.testing.assert_array_equal(
        po.inside_unit_cube(np.array([[0.5, 0.8, 0.2], [0.3, 1.5, 0.9], [0.1, 0.6, -0.1]])),
        [True, False, False])


def test_ess_known_answers():
    # reference tests/test_tools.py:15-50
    assert po.effective_sample_size(np.ones(4)) == pytest.approx(4.0)
    assert po.effective_sample_size(np.array([1.0])) == pytest.approx(1.0)
    assert po.effective_sample_size(np.array([1.0, 0.0, 0.0, 0.0])) == pytest.approx(1.0)


def test_empty_history_and_normalisation():
    # reference tests/test_state_manager.py:204-281
    logw, logz = po.log_weights_and_logz([], [], [])
    assert logw.size == 0 and logz == -np.inf
    rng = np.random.default_rng(0)
    gens = [rng.normal(size=16) for _ in range(3)]
    logw, logz = po.log_weights_and_logz(gens, [0.0, 0.3, 1.0], [0.0, -0.4, -1.0], 1.0)
    assert logw.shape == (48,)
    assert np.exp(logw).sum() == pytest.approx(1.0, abs=1e-5)


def test_first_iteration_and_stay_at_beta_zero():
    # reference tests/test_steps.py:39-56,100-145
    from tempest_b200.registry import Rosenbrock, UniformPrior

    o = po.OraclePS(UniformPrior(-10, 10, 10), Rosenbrock(10), 10, n_particles=32,
                    stream=po.LegacyStream(3))
    o.iterate()
    assert o.cur["beta"] == 0.0 and o.cur["ess"] == 64.0 and o.cur["logz"] == 0.0
    o.iterate()  # N_total = 32 < 64 = target -> stay
    assert o.cur["beta"] == 0.0
    o.iterate()  # N_total = 64 == target -> `<=` keeps beta at 0 (power-of-two N is exact)
    assert o.cur["beta"] == 0.0
    o.iterate()
    assert 0.0 < o.cur["beta"] < 1.0
    assert abs(o.cur["ess"] - 64.0) < 0.64 + 1e-9  # bisection lands within tolerance (:147-200)


def test_systematic_vector_form_equals_loop_form():
    rng = np.random.default_rng(5)
    for n in (1, 7, 64, 1000):
        w = rng.random(n) ** 4
        w /= w.sum()
        for size in (1, 5, 64):
            u0 = rng.random()
            np.testing.assert_array_equal(po.systematic_indices(size, w, u0),
                                          po.systematic_indices_loop(size, w, u0))
    # docstring example of tools.py:201-204 uses 4 draws from [0.6,0.2,0.15,0.05]
    idx = po.systematic_indices(4, np.array([0.6, 0.2, 0.15, 0.05]), 0.5)
    assert idx.min() >= 0 and idx.max() <= 3 and np.all(np.diff(idx) >= 0)


def test_numpy_pairwise_sum_restated():
    rng = np.random.default_rng(1)
    for n in list(range(1, 140)) + [255, 1000, 4097]:
        a = rng.random(n) * rng.choice([1.0, 1e3, 1e-3], n)
        assert po.numpy_pairwise_sum(list(a)) == float(np.sum(a))


def test_student_fit_exits_with_infinite_dof():
    rng = np.random.default_rng(2)
    data = rng.normal(size=(400, 5))
    mu, sigma, nu = po.student_fit(data)
    assert nu == np.inf
    np.testing.assert_allclose(mu, np.median(data, axis=0))
    n = data.shape[0]
    expect = np.cov(data.T) * (n - 1) / n + np.diag(np.var(data, axis=0)) / n
    np.testing.assert_allclose(sigma, expect, rtol=1e-14)
