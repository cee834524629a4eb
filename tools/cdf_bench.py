"""Time tb_cdf_exact on the weight vector of a finished C4 run: the chained single-pass kernel (default) against the
multi-kernel pipeline, both checked bit for bit against numpy's cumsum (run under ncu for the per-kernel split)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import tempest_b200 as tp
from tempest_b200.ensemble import ptr
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
d = 10
s = tp.Sampler(tp.UniformPrior(-10.0, 10.0, d), tp.Rosenbrock(d), d, n_particles=n, vectorize=True, clustering=False,
               random_state=20261018)
s.run(progress=False)
core = s._core
ens, k = core.ensemble, core.k
k.probe(ens, 1.0)
w = k.weights(ens, 1.0, k.probe_out, core.weights_buffer())
k.g_normalize(w, ens.n_total)
ref = np.cumsum(w.cpu().numpy())
peak = 6536.4
for chain in (1, 0):
    k.lib.tb_cdf_set_chain(chain)
    for _ in range(3):
        k.cdf(w, ens.n_total)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        cdf = k.cdf(w, ens.n_total)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    ws = k.ws.bytes("cdf_ws", 64)
    line = (f"{'chained single-pass kernel' if chain else 'multi-kernel pipeline'}: n {ens.n_total} cdf {ms:.4f} ms = "
            f"{16.0 * ens.n_total / ms / 1e6:.0f} GB/s by 16 B/element ({16.0 * ens.n_total / ms / 1e6 / peak:.3f} of {peak} GB/s); "
            f"status {ws[:64].view(torch.int32).cpu().numpy()[:14].tolist()}")
    if chain:
        nt = max(1, int(ws[:4].view(torch.int32)[0].item()))
        off = int(k.lib.tb_cdf_chain_diag_ptr(ptr(ws), ens.n_total)) - ws.data_ptr()
        dg = ws[off: off + 192].view(torch.int64).cpu().numpy().tolist()
        line += (f"; chain diag: multi-round tiles {dg[0]}, rounds in them {dg[1]}, look-back retries {dg[2]}, "
                 f"re-publications {dg[3]}, serial elements {dg[4]}, late prefixes {dg[5]}, look-back rounds {dg[6]}; "
                 f"us per tile: load+aggregate {dg[7] / 1e3 / nt:.2f}, look-back {dg[8] / 1e3 / nt:.2f}, emit {dg[9] / 1e3 / nt:.2f}; "
                 f"failed attempts by reason: not-ready {dg[10]}, tile0 {dg[11]}, all-invalid {dg[12]}, no-offer {dg[13]}, "
                 f"exhausted {dg[14]}, prefix-not-regime {dg[15]}, binade-mismatch {dg[16]}, overflow {dg[17]}")
    print(line)
    print("   bitwise equal to numpy:", np.array_equal(ref.view(np.uint64), cdf.cpu().numpy().view(np.uint64)))
k.lib.tb_cdf_set_chain(0)
torch.cuda.profiler.start()
cdf = k.cdf(w, ens.n_total)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
