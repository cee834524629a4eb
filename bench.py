#!/usr/bin/env python
"""Headline benchmark: Persistent Sampling iterations/s (and logL evaluations/s) on the 10-D
Rosenbrock with 2^20 particles (BASELINE.json configs[3]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--particles P]

A *step* is one PS iteration (reweight -> train -> resample -> mutate -> commit,
tempest/core.py:162-185).  W untimed warm-up iterations (the reference's three beta = 0 prior
generations when W = 3), then exactly K timed iterations bracketed by barrier + synchronize and
timed with CUDA events on the launching stream; a run that terminates inside the timed region is
followed immediately by a fresh run.  Prints ONE JSON line (see the task contract for the keys).

--impl reference times the UNMODIFIED reference (`tempest` from oracle/_ref, a byte-for-byte copy made by
oracle/build_ref.py at build() time; the numpy port oracle/ps_oracle.py only if that copy is missing) on the box's
host cores: the same likelihood / prior / sampler settings at N = --ref-particles (default 2^10, the largest
ensemble whose 25 iterations fit a few minutes), W + K iterations of the reference's own loop body
(core.py:145-146).  `value` is the MEASURED it/s at that N; the O(N) extrapolation to 2^20 is a separate,
labelled field.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "PS iterations/sec + logL evals/sec at N=2^20, 10-D Rosenbrock, 1/2/4/8 B200"
N_DIM = 10
SEED = 20261018


# ----------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def _nvml_loop(self) -> bool:
        """In-process NVML sampling (same counters as nvidia-smi, without forking a process that takes the
        driver lock every 200 ms inside the timed region)."""
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            bits = [("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown")
                     else nv.nvmlClocksThrottleReasonHwSlowdown),
                    ("hw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown",
                                                    getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0))),
                    ("sw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown",
                                                    getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0))),
                    ("sw_power_cap", getattr(nv, "nvmlClocksEventReasonSwPowerCap",
                                             getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0)))]
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                nv.nvmlDeviceGetCurrentClocksThrottleReasons
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        except Exception:
            return False
        while not self._halt.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                mask = get_reasons(h)
                self.rows.append([str(sm), str(mx)] + ["Active" if (mask & b) else "Not Active" for _, b in bits])
            except Exception:
                pass
            self._halt.wait(0.05)
        return True

    def run(self):
        if self._nvml_loop():
            return
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self) -> dict:
        self._halt.set()
        self.join(timeout=2)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i] == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


_REAL_STDOUT = None


def _quiet_stdout() -> None:
    """The contract is ONE JSON line on stdout.  Libraries write banners to fd 1 (NCCL prints its version
    there), so fd 1 is pointed at stderr for the whole run and the JSON line goes to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(text: str) -> None:
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        print(text, flush=True)
    else:
        os.write(_REAL_STDOUT, (text + "\n").encode())


def peak_hbm_gbs() -> tuple:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------
def reference_available() -> bool:
    return os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "tempest"))


def cpu_reference_timing(n_sample: int, n_iter: int, warm: int):
    """`warm` + `n_iter` PS iterations of the CPU implementation at N = n_sample on this host, one process.
    With oracle/_ref present this is the unmodified reference driven through the two calls its own run loop makes
    per iteration (`_not_termination()` + `execute_iteration()`, tempest/core.py:145-146) after
    `_initialize_fresh()`; otherwise the numpy port.  Returns (kind, seconds for the timed iterations, logL calls in
    them, mean MCMC steps per iteration, beta reached)."""
    import numpy as np

    from tempest_b200.registry import Rosenbrock, UniformPrior

    prior, like = UniformPrior(-10.0, 10.0, N_DIM), Rosenbrock(N_DIM)
    if reference_available():
        from oracle.build_ref import import_reference

        tempest = import_reference()
        np.random.seed(SEED % (2**32))          # the reference draws from numpy's global legacy stream
        smp = tempest.Sampler(prior, like, N_DIM, n_particles=n_sample, vectorize=True, clustering=False)
        core = smp._core
        core._initialize_fresh()
        core.n_total = 4096
        state = smp.state

        def iterate():
            core._not_termination()
            core.execute_iteration(None, 0)
            return state.get_current("calls"), state.get_current("steps"), state.get_current("beta")

        kind = "reference"
    else:
        from oracle import ps_oracle as po

        o = po.OraclePS(prior, like, N_DIM, n_particles=n_sample, stream=po.LegacyStream(SEED % (2**32)))
        o.cur.update(iter=0, calls=0, beta=0.0, logz=0.0)
        o.n_total = 4096

        def iterate():
            o.not_terminated()
            o.iterate()
            return o.cur["calls"], o.cur["steps"], o.cur["beta"]

        kind = "port"
    calls0 = 0
    for _ in range(warm):
        calls0, _, _ = iterate()
    steps = []
    t0 = time.perf_counter()
    calls, beta = calls0, 0.0
    for _ in range(n_iter):
        calls, st, beta = iterate()
        steps.append(st)
    dt = time.perf_counter() - t0
    return kind, dt, int(calls - calls0), float(np.mean(steps)), float(beta)


def cpu_kernel_timings(n_total: int = 1 << 22, T: int = 32, n_draw: int = 1 << 20, n_vv: int = 1 << 21) -> dict:
    """Full-size timings of the three CPU kernels SURVEY 8d names, on synthetic data of the shape of a C4/C5 run:
    compute_logw_and_logz (tempest/state_manager.py:418-480), the multinomial resampling call
    (tempest/steps/resample.py:79-82) and volume_variation (tempest/tools.py:58-117).  Reference functions when
    oracle/_ref is present, their numpy restatements otherwise."""
    import numpy as np

    rng = np.random.default_rng(SEED)
    n_gen = n_total // T
    betas = [0.0, 0.0, 0.0] + list(np.geomspace(1e-4, 1.0, T - 3))
    logl = [-0.5 * rng.chisquare(N_DIM, n_gen) * (1.0 + 3.0 * rng.random(n_gen)) for _ in range(T)]
    logz = list(np.linspace(0.0, -30.0, T))
    u = rng.random((n_vv, N_DIM))
    out = {"kind": "port", "shapes": {"n_total": n_total, "T": T, "n_draws": n_draw, "volume_variation_rows": n_vv}}
    if reference_available():
        from oracle.build_ref import import_reference

        tempest = import_reference()
        from tempest.state_manager import StateManager
        from tempest.tools import volume_variation

        sm = StateManager(N_DIM)
        for t in range(T):                      # history through the reference's own commit path
            sm.update_current({"logl": logl[t], "beta": betas[t], "logz": logz[t], "iter": t})
            sm.commit_current_to_history()
        t0 = time.perf_counter()
        logw, _ = sm.compute_logw_and_logz(1.0)
        out["compute_logw_and_logz_s"] = time.perf_counter() - t0
        out["kind"] = "reference"
    else:
        from oracle import ps_oracle as po

        volume_variation = po.volume_variation
        t0 = time.perf_counter()
        logw, _ = po.log_weights_and_logz(logl, betas, logz, 1.0)
        out["compute_logw_and_logz_s"] = time.perf_counter() - t0
    w = np.exp(logw - logw.max())
    w /= w.sum()
    t0 = time.perf_counter()
    np.random.seed(1)
    idx = np.random.choice(np.arange(len(w)), size=n_draw, replace=True, p=w)      # resample.py:79-82
    out["multinomial_resample_s"] = time.perf_counter() - t0
    wv = w[:n_vv] / w[:n_vv].sum()
    t0 = time.perf_counter()
    out["volume_variation_value"] = float(volume_variation(u, wv))
    out["volume_variation_s"] = time.perf_counter() - t0
    out["cells_per_s"] = n_total * T / out["compute_logw_and_logz_s"]
    del idx
    return out


def run_reference_arm(args) -> dict:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        sys.exit(0)
    n_sample = args.ref_particles
    kind, dt, calls, mean_steps, beta = cpu_reference_timing(n_sample, args.steps, args.warmup)
    it_s = args.steps / dt
    what = ("UNMODIFIED reference (tempest 0.2.1 from oracle/_ref, numpy, 1 process)" if kind == "reference"
            else "oracle port of the reference (oracle/_ref missing on this box)")
    sample = (f"{what}: 10-D Rosenbrock, U(-10,10)^10, N={n_sample}, clustering=False, tpCN, multinomial; "
              f"{args.steps} PS iterations after {args.warmup} warm-up = the two calls of the reference's run loop per "
              f"iteration (core.py:145-146), in {dt:.1f} s; {mean_steps:.0f} MCMC steps/iteration; beta reached {beta:.3g}")
    return {
        "impl": "reference", "metric": METRIC, "value": it_s, "unit": "PS iterations/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / it_s, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"10-D Rosenbrock, U(-10,10)^10, clustering=False, tpCN, multinomial (SURVEY C4) at "
                               f"N={n_sample} particles: MEASURED at this N (the reference needs ~1 day at N=2^20)",
                   "n_particles": n_sample},
        "logl_evals_per_s": calls / dt,
        "extrapolated_to_2pow20": {"value": it_s * n_sample / float(1 << 20), "unit": "PS iterations/s",
                                   "how": "O(N) cost model of SURVEY section 6 (an extrapolation, not a measurement)"},
        "cpu_baseline": {"value": it_s, "unit": "PS iterations/s", "cores": 1, "kind": kind, "sample": sample,
                         "logl_evals_per_s": calls / dt, "host_cores": os.cpu_count()},
        "e2e": {"value": it_s, "unit": "PS iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=37)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--particles", type=int, default=1 << 20)
    ap.add_argument("--ref-particles", type=int, default=1024)
    ap.add_argument("--cpu-sample-particles", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-stages", action="store_true")
    args = ap.parse_args()
    _quiet_stdout()

    if args.impl == "reference":
        _emit(json.dumps(run_reference_arm(args)))
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import tempest_b200 as tp
    from tempest_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_particles = args.particles

    stage_acc = {}

    def new_sampler():
        smp = tp.Sampler(tp.UniformPrior(-10.0, 10.0, N_DIM), tp.Rosenbrock(N_DIM), N_DIM,
                         n_particles=n_particles, vectorize=True, clustering=False, random_state=SEED)
        smp._core.stage_ms = stage_acc
        return smp

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- untimed warm-up: one complete run (+ posterior()) of the SAME sampler loads every kernel module (CUDA lazy
    #      loading), sizes the peer-memory channels and fills the allocator pools ("exclude one untimed
    #      warm-up run"); reset() then forgets the history and keeps the buffers, so the timed iterations below do not
    #      pay cudaMalloc (5-10 ms each once peer access is enabled on a multi-GPU box) ---------------------------------
    s = new_sampler()
    core = s._core
    core.stage_ms = {}
    s.run(n_total=4096, progress=False)
    _ = s.posterior()
    del _
    barrier()
    stage_acc.clear()

    # ---- device-timed region: K PS iterations after W warm-up iterations ---------------------------
    core.reset()                               # same sampler, same buffers, history cleared
    core.stage_ms = stage_acc
    core.profile = args.profile_stages
    core.kernel_timing = None
    core.n_total = 4096                       # run() default n_total (sampler.py:165)
    runs_T = []

    def one_iteration():
        nonlocal s, core
        if not core._not_termination():
            runs_T.append(core.state.get_history_length())
            core.reset()                       # next run: same sampler, same buffers, history cleared
            core.n_total = 4096
        core.execute_iteration(export=False)

    for _ in range(args.warmup):
        one_iteration()
    barrier()
    import gc

    gc.collect()
    gc.disable()                               # no collector pauses inside the timed regions
    core.kernel_timing = {}                    # CUDA events around every launch of the two hot kernels (timed region only)
    clocks = ClockSampler(local)
    clocks.start()
    calls0 = core.state.raw("calls")
    lib = _lib.load()
    launches0 = _lib.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    calls_acc = 0
    trace = os.environ.get("BENCH_TRACE")
    marks = []
    it_events = []                             # one event per iteration boundary (no synchronisation)
    for _ in range(args.steps):
        before = core.state.raw("calls")
        one_iteration()
        after = core.state.raw("calls")
        calls_acc += after - (before if after >= before else 0)
        ev_it = torch.cuda.Event(enable_timing=True)
        ev_it.record()
        it_events.append(ev_it)
        if trace:
            marks.append(round(time.perf_counter(), 4))
    ev1.record()
    if trace:
        print("iteration host marks (s):", [round(b - a, 4) for a, b in zip(marks[:-1], marks[1:])], file=sys.stderr)
    launches = _lib.launch_count - launches0
    barrier()
    ms = ev0.elapsed_time(ev1)
    iteration_ms = [round(a.elapsed_time(b), 3) for a, b in zip([ev0] + it_events[:-1], it_events)]
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clock_info = clocks.stop()
    hot = {name: (sum(a.elapsed_time(b) for a, b, _ in rows), sum(w for _, _, w in rows), len(rows))
           for name, rows in core.kernel_timing.items()}
    core.kernel_timing = None
    it_s = args.steps / (ms * 1e-3)
    evals_s = calls_acc / (ms * 1e-3)
    stage_ms = dict(stage_acc)

    # ---- end to end through the public API (host in, host out) -----------------------------------------
    del core, s                                # (sharded runs: hands the pooled peer-memory buffers back, as a user's
    gc.collect()                               #  previous Sampler going out of scope would)
    barrier()
    t0 = time.perf_counter()
    s2 = new_sampler()
    s2.run(n_total=4096, progress=False)
    t_run = time.perf_counter() - t0
    logz, _ = s2.evidence()
    x, w, l = s2.posterior()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if trace:
        print(f"e2e: run {t_run:.3f} s, posterior {dt - t_run:.3f} s", file=sys.stderr)
    T2 = s2.state.get_history_length()
    d2h = (x.nbytes + w.nbytes + l.nbytes) / T2 + 16 * 8 * 12 + 3 * 2048 * 8
    h2d = 3 * T2 * 8 + 64
    e2e = {"value": T2 / dt, "unit": "PS iterations/s", "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "iterations": T2, "seconds": dt, "logz": float(logz),
           "logl_evals_per_s": s2.state.get_current("calls") / dt, "posterior_samples": int(x.shape[0]),
           "note": "Sampler(...).run(4096) + evidence() + posterior() to host numpy; inputs are the problem "
                   "definition (parameters), outputs the weighted posterior sample"}
    del x, w, l
    gc.enable()

    # ---- rooflines.  Lead: the time-dominant kernel, the persistent Metropolis kernel (one launch per mutation),
    #      timed live with CUDA events around every launch inside the timed region above.  Algorithmic work per
    #      walker-step at D = 10 (SURVEY 8d flop model): D(D+1) [L z] + 2 D^2 [quadratic form of the proposal; the
    #      current one is cached] + 6 D [Rosenbrock] + ~30 [exp, 2 log, Student-t scale] = 400 fp64 flop; it moves
    #      2 (D+3) 8 B = 208 B of HBM per walker-step (0.5 flop/B below the fp64 ridge: fp64-pipe / issue bound).
    #      Denominator: fp64 FMA peak measured here with tb_fp64_peak_run (register-resident DFMA chains). ----------
    from tempest_b200.steps import Kernels

    core2 = s2._core
    ens = core2.ensemble
    n_hist = ens.n_total

    def cuda_ms(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(reps):
            fn(i)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    peak_out = torch.zeros(1, dtype=torch.float64, device="cuda")
    iters = 1 << 16
    from tempest_b200.ensemble import ptr, stream_ptr
    fp64_ms = min(cuda_ms(lambda i=0: lib.tb_fp64_peak_run(iters, ptr(peak_out), stream_ptr()), reps=3) for _ in range(3))
    fp64_peak = lib.tb_fp64_peak_flops(iters) / (fp64_ms * 1e-3) / 1e12
    flop_per_walker_step = 400.0
    m_ms, m_work, m_n = hot.get("mcmc", (0.0, 0.0, 0))
    achieved_tf = flop_per_walker_step * m_work / (m_ms * 1e-3) / 1e12 if m_ms > 0 else 0.0
    roofline = {"bound": "fp64", "kernel": "mcmc_run_fast<10, tpCN, Philox, single mode, Rosenbrock> (tb_mcmc_fast.cuh): "
                                           "persistent cooperative kernel, one launch per mutation",
                "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved_tf / fp64_peak,
                "traffic": None, "flop_per_walker_step": flop_per_walker_step,
                "launches": m_n, "launch_ms": m_ms / max(m_n, 1), "walker_steps_per_launch": m_work / max(m_n, 1),
                "ms_per_step": (m_ms / (m_work / n_particles * world)) if m_work else None,
                "share_of_timed_region": m_ms / ms if ms else None,
                "peak_source": f"measured here: tb_fp64_peak_run, {iters} x 8 DFMA chains per thread, "
                               f"{fp64_ms:.3f} ms (nominal B200 fp64: 37-40 TFLOP/s)"}
    peak, peak_src = peak_hbm_gbs()
    others = {}
    nb_ms, nb_work, nb_n = hot.get("next_beta", (0.0, 0.0, 0))
    if nb_ms > 0:
        others["next_beta_kernel (in situ, all launches of the timed region)"] = {
            "ms": nb_ms / nb_n, "launches": nb_n, "bytes_per_particle_pass": 16,
            "gbs": 16.0 * nb_work / nb_ms / 1e6, "share_of_timed_region": nb_ms / ms}
    probe_ms = cuda_ms(lambda i=0: Kernels.probe(core2.k, ens, 0.3 + 0.01 * i))     # local kernel only
    others["probe_kernel"] = {"ms": probe_ms, "bytes_per_particle": 16, "gbs": 16.0 * n_hist / probe_ms / 1e6,
                              "particles": n_hist}
    prof = os.path.join(ROOT, "profiles", "r01_probe.json")
    if os.path.exists(prof):      # dram bytes per particle from the committed `ncu --set full` capture, scaled to this N_total
        pj = json.load(open(prof))
        others["probe_kernel"]["traffic"] = pj["traffic_bytes_per_particle"] * n_hist
    wbuf = core2.weights_buffer()
    Kernels.probe(core2.k, ens, 1.0)
    ms_w = cuda_ms(lambda i=0: core2.k.weights(ens, 1.0, core2.k.probe_out, wbuf))
    others["weights_kernel"] = {"ms": ms_w, "bytes_per_particle": 24, "gbs": 24.0 * n_hist / ms_w / 1e6}
    ms_c = cuda_ms(lambda i=0: core2.k.cdf(wbuf, n_hist), reps=10)
    others["cdf_exact"] = {"ms": ms_c, "bytes_per_particle": 16, "gbs": 16.0 * n_hist / ms_c / 1e6,
                           "note": "multi-kernel pipeline (default); algorithmic bytes: read w, write cdf (SURVEY 8d)"}
    if world == 1:
        lib.tb_cdf_set_chain(1)
        ms_cc = cuda_ms(lambda i=0: core2.k.cdf(wbuf, n_hist), reps=10)
        lib.tb_cdf_set_chain(0)
        others["cdf_exact (chained look-back kernel, option)"] = {
            "ms": ms_cc, "bytes_per_particle": 16, "gbs": 16.0 * n_hist / ms_cc / 1e6,
            "note": "guess pass + one chained kernel; stalls at the ~55 multi-binade jumps of the running sum (DESIGN 4)"}
    if world == 1:
        ms_n = cuda_ms(lambda i=0: core2.k.next_beta(ens, 0.5, 2.0 * n_particles, 0), reps=5)
        npr = float(core2.k.ws.f64("nb_res", 16)[6].item())
        others["next_beta_kernel (isolated)"] = {"ms": ms_n, "probes": npr, "bytes_per_particle": 16 * npr,
                                                 "gbs": 16.0 * npr * n_hist / ms_n / 1e6}
    for v in others.values():
        v["frac"] = v["gbs"] / peak
    roofline["other_kernels"] = others
    roofline["hbm_peak_gbs"] = peak
    roofline["hbm_peak_source"] = peak_src

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only): a bounded sample of the same workload ----------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_s = args.cpu_sample_particles
        k_s, w_s = min(args.steps, 15), min(args.warmup, 5)
        kind, dt, calls, mean_steps, beta = cpu_reference_timing(n_s, k_s, w_s)
        cpu_it = k_s / dt
        what = ("UNMODIFIED reference (tempest from oracle/_ref, numpy, 1 process)" if kind == "reference"
                else "oracle port of the reference (oracle/_ref missing)")
        cpu = {"value": cpu_it, "unit": "PS iterations/s", "cores": 1, "kind": kind,
               "sample": f"{what}, same workload at N={n_s}: {k_s} PS iterations after {w_s} warm-up in {dt:.1f} s, "
                         f"{mean_steps:.0f} MCMC steps/iteration; value is MEASURED at N={n_s}",
               "logl_evals_per_s": calls / dt, "host_cores": os.cpu_count(),
               "extrapolated_to_bench_N": {"value": cpu_it * n_s / float(n_particles), "unit": "PS iterations/s",
                                           "how": "O(N) cost model (SURVEY section 6); an extrapolation"}}
        try:
            cpu["kernels_full_size"] = cpu_kernel_timings()
        except MemoryError as exc:     # pragma: no cover - host too small for the 2 x 1 GiB temporaries
            cpu["kernels_full_size"] = {"skipped": str(exc)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": it_s, "unit": "PS iterations/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"10-D Rosenbrock, U(-10,10)^10, N={n_particles} particles, clustering=False, "
                                   "tpCN, multinomial resampling, ess_ratio=2, n_total=4096 (SURVEY C4)",
                       "l2": "inputs larger than L2 (history >= 3 generations x 92 MB); no flush",
                       "rng": f"Philox4x32-10 seed {SEED}; normals: Box-Muller on 32-bit words with fp32 radius / angle "
                              "promoted to fp64 (as curand_normal; A/B against fp64 Box-Muller in "
                              "profiles/r02_normals_ab.txt), gamma / accept tests in fp64",
                       "timed_window": f"PS iterations {args.warmup + 1}..{args.warmup + args.steps} of the beta ladder "
                                       "(a run has ~36; a finished run is followed by a fresh one); `e2e` is a whole run "
                                       "incl. the cheap warm-up and the expensive beta -> 1 iterations + posterior()",
                       "runs_completed_T": runs_T},
            "logl_evals_per_s": evals_s, "gpu_launches": None, "clocks": clock_info, "e2e": e2e,
            "roofline": roofline, "cpu_baseline": cpu, "stage_ms": stage_ms, "iteration_ms": iteration_ms,
        }
        line["gpu_launches"] = int(launches)
        _emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
