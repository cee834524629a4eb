"""CPU-only checks: host-side scalar logic, configuration, and that the C-ABI library loads and
exports every symbol include/tempest_b200.h declares (no compute calls without a GPU)."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def test_library_exports_every_declared_symbol():
    from tempest_b200 import _lib, build

    build.build(verbose=False)
    lib = _lib.load_library()
    header = open(os.path.join(ROOT, "include", "tempest_b200.h")).read()
    declared = set(re.findall(r"\b(tb_[a-z0-9_]+)\s*\(", header)) - {"tb_stream_t"}
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert set(_lib.SIGNATURES) <= declared
    assert lib.tb_version() == 100


def test_uniform_weights_ess_matches_numpy():
    from tempest_b200.steps import uniform_weights_ess, numpy_pairwise_sum_equal

    for n in list(range(1, 300)) + [1000, 1023, 1024, 3000, 4097, 12345, 65536, 3 * 21, 1 << 20]:
        w = np.ones(n)
        w = w / np.sum(w)
        assert uniform_weights_ess(n) == 1.0 / np.sum(w**2.0), n
    for n in (5, 77, 129, 5000):
        assert numpy_pairwise_sum_equal(0.1, n) == float(np.sum(np.full(n, 0.1)))


def test_percentile_restatement_matches_numpy():
    from tempest_b200.steps import numpy_lerp, percentile_position

    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 10, 101, 1000, 4099):
        w = rng.random(n) ** 5
        s = np.sort(w)
        for p in np.linspace(0, 99, 1000)[:: max(1, 1000 // 97)]:
            lo, hi, g = percentile_position(n, float(p))
            assert numpy_lerp(s[lo], s[hi], g) == np.percentile(w, p), (n, p)


def test_config_defaults_and_validation():
    from tempest_b200.config import SamplerConfig

    f = lambda x: x  # noqa: E731
    c = SamplerConfig(prior_transform=f, log_likelihood=f, n_dim=3)
    assert (c.n_particles, c.n_steps, c.n_max_steps, c.resample, c.sample) == (6, 1, 20, "mult", "tpcn")
    assert str(c.output_dir) == "states" and c.output_label == "ps"
    with pytest.raises(ValueError, match="Invalid sampler 'hmc'"):
        SamplerConfig(prior_transform=f, log_likelihood=f, n_dim=3, sample="hmc")
    with pytest.raises(ValueError, match="Cannot vectorize likelihood with blobs"):
        SamplerConfig(prior_transform=f, log_likelihood=f, n_dim=3, vectorize=True, blobs_dtype="f8")
    with pytest.raises(ValueError, match="both periodic and reflective"):
        SamplerConfig(prior_transform=f, log_likelihood=f, n_dim=3, periodic=[0], reflective=[0])
    with pytest.raises(ValueError, match="n_dim must be int"):
        SamplerConfig(prior_transform=f, log_likelihood=f, n_dim=2.5)


def test_product_fails_loudly_without_cuda():
    import torch
    import tempest_b200 as tp

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="CUDA"):
        tp.Sampler(tp.UniformPrior(-1, 1, 2), tp.Rosenbrock(2), 2, vectorize=True, clustering=False)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tempest_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("# oracle", ""), fn


def test_registry_matches_readme_rosenbrock():
    from tempest_b200.registry import Rosenbrock, UniformPrior

    rng = np.random.default_rng(3)
    x = UniformPrior(-10, 10, 10)(rng.random((257, 10)))
    readme = -np.sum(10.0 * (x[:, ::2] ** 2.0 - x[:, 1::2]) ** 2.0 + (x[:, ::2] - 1.0) ** 2.0, axis=1)
    np.testing.assert_array_equal(Rosenbrock(10)(x), readme)
    np.testing.assert_array_equal(UniformPrior(-10, 10, 10)(np.full(10, 0.25)), 20 * np.full(10, 0.25) - 10)


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    from oracle.philox import philox4x32

    kat = [((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
           ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
           ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
            (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1))]
    for ctr, key, want in kat:
        got = philox4x32(*ctr, *key)
        assert tuple(int(v) for v in got) == want


def test_callable_bridge_decides_batched_vs_rowwise_prior_on_the_host():
    """callables.py: a prior that is elementwise may be called on the whole batch, one that indexes
    coordinates must be called row by row like the reference does (mcmc.py:157)."""
    import torch

    from tempest_b200.callables import CallableBridge
    from tempest_b200.config import SamplerConfig
    from tempest_b200.registry import Rosenbrock, UniformPrior

    def elementwise(u):
        return 20.0 * u - 10.0

    def rowwise(u):
        x = np.empty(4)
        x[0], x[1], x[2], x[3] = u[0], 2 * u[1], 3 * u[2], 4 * u[3]
        return x

    def like(x, shift=0.0):
        return -np.sum(np.atleast_2d(x) ** 2, axis=1) + shift

    dev = torch.device("cpu")
    b1 = CallableBridge(SamplerConfig(elementwise, like, 4, vectorize=True), dev)
    assert b1.external and b1.prior_batched and not b1.like_registry
    b2 = CallableBridge(SamplerConfig(rowwise, like, 4, vectorize=True), dev)
    assert b2.external and not b2.prior_batched
    b3 = CallableBridge(SamplerConfig(UniformPrior(-1, 1, 4), Rosenbrock(4), 4, vectorize=True), dev)
    assert not b3.external and b3.prior_registry and b3.like_registry
    u = torch.rand((5, 4), dtype=torch.float64)
    x2 = b2.prior(u, None)
    np.testing.assert_array_equal(x2.numpy(), u.numpy() * np.array([1.0, 2.0, 3.0, 4.0]))
    np.testing.assert_array_equal(b1.like(b1.prior(u, None)).numpy(), like(elementwise(u.numpy())))
    # vectorize=False: one sample per call (core.py:323-326)
    seen = []

    def one(x):
        seen.append(x.shape)
        return float(-np.sum(x * x))

    b4 = CallableBridge(SamplerConfig(elementwise, one, 4, vectorize=False), dev)
    out = b4.like(torch.ones((3, 4), dtype=torch.float64))
    assert seen == [(4,), (4,), (4,)] and out.shape == (3,)


def test_vectorised_percentile_positions_equal_the_scalar_form():
    from tempest_b200.steps import percentile_position, percentile_positions

    grid = np.linspace(0, 99, 1000)
    for n in (1, 2, 5, 20, 4097, 123457, 37748736):
        lo, hi, g = percentile_positions(n, grid)
        ref = [percentile_position(n, float(p)) for p in grid]
        np.testing.assert_array_equal(lo, [r[0] for r in ref])
        np.testing.assert_array_equal(hi, [r[1] for r in ref])
        np.testing.assert_array_equal(g, [r[2] for r in ref])


def test_plain_c_caller_compiles_and_links_against_the_abi(tmp_path):
    """examples/next_beta_c_abi.c uses the library from C (no Python, no torch): it must compile as C
    against include/tempest_b200.h and link against the shared object."""
    import shutil
    import subprocess

    from tempest_b200 import _lib

    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc) or not os.path.exists(_lib.LIB_PATH):
        pytest.skip("nvcc or the built library is not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "next_beta_demo"
    cmd = [nvcc, "-x", "c", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "next_beta_c_abi.c"),
           "-L", os.path.dirname(_lib.LIB_PATH), "-ltempest_b200", "-lcudart", "-o", str(out)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert out.exists()


def test_symmetric_pool_leases_and_returns_objects():
    """sharded._pool_acquire / _pool_release: an idle object that fits is reused, otherwise the factory runs."""
    from tempest_b200 import sharded

    sharded._SYMM_POOL.pop(("t",), None)
    made = []

    def factory():
        made.append(object())
        return {"cap": 10 * len(made), "id": made[-1]}

    a = sharded._pool_acquire(("t",), factory)
    leased = [(("t",), a)]
    sharded._pool_release(leased)
    assert leased == [] and len(made) == 1
    assert sharded._pool_acquire(("t",), factory, fits=lambda o: o["cap"] >= 5) is a          # reused
    sharded._pool_release([(("t",), a)])
    b = sharded._pool_acquire(("t",), factory, fits=lambda o: o["cap"] >= 15)                 # too small: a new one
    assert b is not a and len(made) == 2
    sharded._SYMM_POOL.pop(("t",), None)


def test_bench_reference_arm_contract_on_cpu():
    """`bench.py --impl reference` runs the UNMODIFIED reference (oracle/_ref) on the host and prints ONE JSON line with
    the contract keys; a non-zero rank exits without work.  Tiny N so it takes seconds."""
    import json
    import subprocess
    import sys

    from conftest import ROOT

    if not os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "tempest")):
        pytest.skip("oracle/_ref not built here (python oracle/build_ref.py needs /root/reference)")
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
           "--ref-particles", "64"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["higher_is_better"] is True and j["unit"] == "PS iterations/s"
    assert j["value"] > 0 and j["steps"] == 2 and j["warmup"] == 1
    assert j["cpu_baseline"]["kind"] == "reference" and j["cpu_baseline"]["cores"] == 1
    assert j["e2e"]["value"] == j["value"] and j["e2e"]["h2d_bytes_per_step"] == 0
    assert j["config"]["n_particles"] == 64 and "extrapolated_to_2pow20" in j
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=60, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
