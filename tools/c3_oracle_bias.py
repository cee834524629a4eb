"""CPU oracle (bit-exact restatement of the reference) on the C3 likelihood at small N, run to beta = 1:
does the reference algorithm itself over-estimate logZ at d = 50 (analytic -149.787)?
    python tools/c3_oracle_bias.py SEED [N]      -> appends one line to gpurun_out/c3_oracle_bias.txt"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import ps_oracle as po
from tempest_b200.registry import GaussianLikelihood, UniformPrior
seed = int(sys.argv[1]); n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
d = 50
t0 = time.time()
o = po.OraclePS(UniformPrior(-10.0, 10.0, d), GaussianLikelihood.ar1(d, 0.5), d, n_particles=n, stream=po.LegacyStream(seed))
o.run(4096)
line = (f"oracle d=50 AR(1) Gaussian N={n} seed={seed}: T={len(o.hist['beta'])} logZ={o.evidence()[0]:.4f} "
        f"(analytic {-50 * np.log(20.0):.4f}) mean steps {np.mean(o.hist['steps']):.0f} in {time.time() - t0:.0f} s")
print(line)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "c3_oracle_bias.txt"), "a").write(line + "\n")
