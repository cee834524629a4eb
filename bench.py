#!/usr/bin/env python
"""Headline benchmark: Persistent Sampling iterations/s (and logL evaluations/s) on the 10-D
Rosenbrock with 2^20 particles (BASELINE.json configs[3]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--particles P]

A *step* is one PS iteration (reweight -> train -> resample -> mutate -> commit,
tempest/core.py:162-185).  W untimed warm-up iterations (the reference's three beta = 0 prior
generations when W = 3), then exactly K timed iterations bracketed by barrier + synchronize and
timed with CUDA events on the launching stream; a run that terminates inside the timed region is
followed immediately by a fresh run.  Prints ONE JSON line (see the task contract for the keys).

--impl reference times the CPU restatement of the reference (oracle/ps_oracle.py; the reference
is pure Python/numpy and cannot travel to the GPU box) on a bounded sample of the same workload.
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
def cpu_reference_timing(n_sample: int, n_iter: int, warm: int):
    """Oracle port of the reference on one host core.  Runs PS iterations of a fresh run (beta ladder from
    0 towards 1, restarting if the run completes) and returns (seconds for `n_iter` iterations after `warm`,
    logL calls in them, mean MCMC steps per iteration)."""
    import numpy as np

    from oracle import ps_oracle as po
    from tempest_b200.registry import Rosenbrock, UniformPrior

    def fresh():
        o = po.OraclePS(UniformPrior(-10.0, 10.0, N_DIM), Rosenbrock(N_DIM), N_DIM, n_particles=n_sample,
                        stream=po.LegacyStream(SEED % (2**32)))
        o.cur.update(iter=0, calls=0, beta=0.0, logz=0.0)
        o.n_total = 1           # like the GPU workload: the run ends when beta reaches 1
        return o

    o = fresh()
    for _ in range(warm):
        o.iterate()
    calls, steps = 0, []
    t0 = time.perf_counter()
    for _ in range(n_iter):
        if not o.not_terminated():
            o = fresh()
        c0 = o.cur["calls"]
        o.iterate()
        calls += o.cur["calls"] - c0
        steps.append(o.cur["steps"])
    dt = time.perf_counter() - t0
    return dt, calls, float(np.mean(steps))


def run_reference_arm(args) -> dict:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        sys.exit(0)
    n_sample = args.ref_particles
    dt, calls, mean_steps = cpu_reference_timing(n_sample, args.steps, args.warmup)
    scale = n_sample / float(args.particles)          # PS iteration cost is linear in N
    it_s = args.steps / dt * scale
    sample = (f"oracle port of the reference (numpy, 1 thread), 10-D Rosenbrock, N={n_sample}: {args.steps} PS "
              f"iterations after {args.warmup} warm-up in {dt:.1f} s ({mean_steps:.0f} MCMC steps/iteration); "
              f"it/s scaled by N_sample/2^20 (O(N) cost model, extrapolation)")
    return {
        "impl": "reference", "metric": METRIC, "value": it_s, "unit": "PS iterations/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / it_s, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"10-D Rosenbrock, U(-10,10)^10, N=2^20 particles, clustering=False, tpCN, "
                               f"multinomial (SURVEY C4); CPU arm measured at N={n_sample}"},
        "logl_evals_per_s": calls / dt,
        "cpu_baseline": {"value": it_s, "unit": "PS iterations/s", "cores": 1, "kind": "port", "sample": sample},
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
    ap.add_argument("--ref-particles", type=int, default=256)
    ap.add_argument("--cpu-sample-particles", type=int, default=96)
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

    # ---- untimed library warm-up: one complete run loads every kernel module (CUDA lazy loading), sizes the
    #      NCCL channels and fills the allocator pools ("exclude one untimed warm-up run") ----------
    ws_ = tp.Sampler(tp.UniformPrior(-10.0, 10.0, N_DIM), tp.Rosenbrock(N_DIM), N_DIM, n_particles=n_particles,
                     vectorize=True, clustering=False, random_state=SEED)
    ws_.run(n_total=4096, progress=False)      # same shape as the timed workload: the caching allocator keeps its blocks
    _ = ws_.posterior()                        # ... and the page-locked staging blocks of the posterior read-back
    del ws_, _
    barrier()

    # ---- device-timed region: K PS iterations after W warm-up iterations ---------------------------
    s = new_sampler()
    core = s._core
    core.profile = args.profile_stages
    core._initialize_fresh()
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
    for _ in range(args.steps):
        before = core.state.raw("calls")
        one_iteration()
        after = core.state.raw("calls")
        calls_acc += after - (before if after >= before else 0)
        if trace:
            marks.append(round(time.perf_counter(), 4))
    ev1.record()
    if trace:
        print("iteration host marks (s):", [round(b - a, 4) for a, b in zip(marks[:-1], marks[1:])], file=sys.stderr)
    launches = _lib.launch_count - launches0
    barrier()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clock_info = clocks.stop()
    it_s = args.steps / (ms * 1e-3)
    evals_s = calls_acc / (ms * 1e-3)
    stage_ms = dict(stage_acc)

    # ---- end to end through the public API (host in, host out) -----------------------------------------
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

    # ---- roofline of the HBM-bound hot kernel, measured live on the FULL persistent ensemble of that run:
    #      the ESS probe streams logl[] and C[] once = 16 algorithmic bytes per stored particle ---------------
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

    peak, peak_src = peak_hbm_gbs()
    probe_ms = cuda_ms(lambda i=0: Kernels.probe(core2.k, ens, 0.3 + 0.01 * i))     # local kernel only
    achieved = 16.0 * n_hist / (probe_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "probe_kernel (tb_reweight.cu; same loop as next_beta_kernel)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "algorithmic_bytes_per_particle": 16, "particles": n_hist, "launch_ms": probe_ms,
                "peak_source": peak_src}
    prof = os.path.join(ROOT, "profiles", "r01_probe.json")
    if os.path.exists(prof):      # dram bytes per particle from the committed `ncu --set full` capture, scaled to this N_total
        pj = json.load(open(prof))
        roofline["traffic"] = pj["traffic_bytes_per_particle"] * n_hist
        roofline["traffic_source"] = (f"{pj['source']}: dram read+write {pj['dram_bytes_read'] + pj['dram_bytes_write']} B "
                                      f"at N_total={pj['n_total']} (algorithmic {pj['algorithmic_bytes']} B), scaled by N_total")
    wbuf = core2.weights_buffer()
    Kernels.probe(core2.k, ens, 1.0)
    others = {}
    ms_w = cuda_ms(lambda i=0: core2.k.weights(ens, 1.0, core2.k.probe_out, wbuf))
    others["weights_kernel"] = {"ms": ms_w, "bytes_per_particle": 24, "gbs": 24.0 * n_hist / ms_w / 1e6}
    ms_c = cuda_ms(lambda i=0: core2.k.cdf(wbuf, n_hist), reps=10)
    others["cdf_exact (6 kernels)"] = {"ms": ms_c, "bytes_per_particle": 32, "gbs": 32.0 * n_hist / ms_c / 1e6}
    if world == 1:
        ms_n = cuda_ms(lambda i=0: core2.k.next_beta(ens, 0.5, 2.0 * n_particles, 0), reps=5)
        npr = float(core2.k.ws.f64("nb_res", 16)[6].item())
        others["next_beta_kernel"] = {"ms": ms_n, "probes": npr, "bytes_per_particle": 16 * npr,
                                      "gbs": 16.0 * npr * n_hist / ms_n / 1e6}
    for v in others.values():
        v["frac"] = v["gbs"] / peak
    roofline["other_kernels"] = others

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only) ---------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_s = args.cpu_sample_particles
        dt, calls, mean_steps = cpu_reference_timing(n_s, args.steps, args.warmup)
        cpu_it = args.steps / dt * (n_s / float(n_particles))
        cpu = {"value": cpu_it, "unit": "PS iterations/s", "cores": 1, "kind": "port",
               "sample": f"oracle port (numpy restatement of the reference, 1 thread), same workload at N={n_s}: "
                         f"{args.steps} PS iterations after {args.warmup} warm-up in {dt:.1f} s, {mean_steps:.0f} MCMC "
                         f"steps/iteration, {calls / dt:.0f} logL evals/s; it/s scaled by N_sample/N (O(N) cost "
                         f"model, an extrapolation) to N={n_particles}",
               "logl_evals_per_s": calls / dt, "host_cores": os.cpu_count()}

    if rank == 0:
        line = {
            "metric": METRIC, "value": it_s, "unit": "PS iterations/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"10-D Rosenbrock, U(-10,10)^10, N={n_particles} particles, clustering=False, "
                                   "tpCN, multinomial resampling, ess_ratio=2, n_total=4096 (SURVEY C4)",
                       "l2": "inputs larger than L2 (history >= 3 generations x 92 MB); no flush",
                       "rng": f"Philox4x32-10 seed {SEED}", "runs_completed_T": runs_T},
            "logl_evals_per_s": evals_s, "gpu_launches": None, "clocks": clock_info, "e2e": e2e,
            "roofline": roofline, "cpu_baseline": cpu, "stage_ms": stage_ms,
        }
        line["gpu_launches"] = int(launches)
        _emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
