"""A/B of the production normal variates: fp32 Box-Muller (MUFU log / sin / cos on 32-bit words, promoted to fp64;
the default build) against an fp64 Box-Muller build (-DTB_NORMALS_F64) on the same workload and seeds.

    python tools/ab_normals.py [n_particles] [n_seeds]        (builds the second library next to the first)

Reports logZ, posterior means / variances and the Metropolis step time of both builds; a systematic effect of the
24-bit normals would show as a shift larger than the seed-to-seed scatter."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
n = int(sys.argv[1]) if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else 1 << 16
seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 6

if "--child" in sys.argv:
    import numpy as np, torch, time
    import tempest_b200 as tp
    d = 10
    out = []
    for seed in range(seeds):
        s = tp.Sampler(tp.UniformPrior(-10.0, 10.0, d), tp.Rosenbrock(d), d, n_particles=n, vectorize=True,
                       clustering=False, random_state=1000 + seed)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        s.run(progress=False)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        x, w, l = s.posterior()
        m = np.average(x, weights=w, axis=0)
        v = np.average((x - m) ** 2, weights=w, axis=0)
        out.append(dict(seed=seed, logz=s.evidence()[0], T=s.state.get_history_length(), seconds=dt,
                        steps=int(np.sum(s.state.get_history("steps"))), mean0=m[0], mean1=m[1], var0=v[0], var1=v[1]))
    print("RESULT " + json.dumps(out))
    sys.exit(0)

env = dict(os.environ, TB_OBJ_SUFFIX="_f64n", TB_LIB_NAME="libtempest_b200_f64n.so", TB_NVCC_EXTRA="-DTB_NORMALS_F64")
if not os.path.exists(os.path.join(ROOT, "tempest_b200", "lib", "libtempest_b200_f64n.so")):
    subprocess.run([sys.executable, "-m", "tempest_b200.build"], env=env, check=True, cwd=ROOT, stdout=subprocess.DEVNULL)
import numpy as np
res = {}
for name, lib in (("fp32 Box-Muller (default)", "libtempest_b200.so"), ("fp64 Box-Muller", "libtempest_b200_f64n.so")):
    e = dict(os.environ, TEMPEST_B200_LIB=os.path.join(ROOT, "tempest_b200", "lib", lib))
    p = subprocess.run([sys.executable, __file__, str(n), str(seeds), "--child"], env=e, capture_output=True, text=True)
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("RESULT ")]
    assert line, p.stdout[-2000:] + p.stderr[-2000:]
    res[name] = json.loads(line[0][7:])
print(f"# A/B of the production normals: 10-D Rosenbrock, N = {n}, {seeds} seeds per build (tools/ab_normals.py)")
for name, rows in res.items():
    z = np.array([r["logz"] for r in rows])
    print(f"{name}: logZ {z.mean():.4f} +- {z.std(ddof=1) / np.sqrt(len(z)):.4f} (scatter {z.std(ddof=1):.4f}); "
          f"E[x0] {np.mean([r['mean0'] for r in rows]):.4f}, E[x1] {np.mean([r['mean1'] for r in rows]):.4f}, "
          f"Var[x0] {np.mean([r['var0'] for r in rows]):.4f}, Var[x1] {np.mean([r['var1'] for r in rows]):.4f}; "
          f"run {np.mean([r['seconds'] for r in rows][1:]):.3f} s, {np.mean([r['steps'] for r in rows]):.0f} steps")
