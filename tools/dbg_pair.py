"""Debug: first divergence between the device run and the oracle on tapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import tempest_b200 as tp
from oracle import ps_oracle as po
from oracle.gen_golden import cases
from tempest_b200.rng import TapeSource
name = sys.argv[1] if len(sys.argv) > 1 else "rosen10_n64_tpcn_mult"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
prior, like, kw, n_total, seed = cases()[name]
o = po.OraclePS(prior, like, stream=po.LegacyStream(seed), record=True, **kw)
o.run(n_total, max_iterations=iters)
s = tp.Sampler(prior, like, vectorize=True, **kw)
core = s._core
core.rng = TapeSource(o.tapes, core.device)
core._initialize_fresh(); core.n_total = int(n_total)
for t in range(iters):
    core.execute_iteration()
    st = s.state
    u = st.get_current("u"); l = st.get_current("logl")
    du = np.abs(u - np.array(o.hist["u"][t])).max()
    print(t, "beta", st.raw("beta"), o.hist["beta"][t], "steps", st.raw("steps"), o.hist["steps"][t],
          "acc", st.raw("acceptance"), o.hist["acceptance"][t], "eff", st.raw("efficiency"), o.hist["efficiency"][t], "max|du|", du,
          "logz", st.raw("logz"), o.hist["logz"][t])
