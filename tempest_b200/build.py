"""Builds libtempest_b200.so in-tree with nvcc for sm_100a (no torch headers involved).

    python -m tempest_b200.build            # incremental (per-source objects, compiled in parallel)
    python -m tempest_b200.build --force
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build" + os.environ.get("TB_OBJ_SUFFIX", ""))
LIB = os.path.join(LIBDIR, os.environ.get("TB_LIB_NAME", "libtempest_b200.so"))   # TB_LIB_NAME / TB_NVCC_EXTRA: A/B builds
SOURCES = ["tb_reweight.cu", "tb_resample.cu", "tb_cdf.cu", "tb_xcoll.cu", "tb_moments.cu", "tb_linalg.cu", "tb_mcmc.cu", "tb_cluster.cu",
           "tb_mcmc_fast_a.cu", "tb_mcmc_fast_b.cu", "tb_mcmc_fast_c.cu", "tb_mcmc_fast_d.cu", "tb_mcmc_fast_e.cu",
           "tb_mcmc_fast_f.cu", "tb_mcmc_wide.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-warn-spills",
] + os.environ.get("TB_NVCC_EXTRA", "").split()


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return "nvcc"


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(os.path.dirname(HERE), "include", "tempest_b200.h"))
    return hs


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build() -> bool:
    return _stale(LIB, [os.path.join(CSRC, s) for s in SOURCES] + _headers())


def build(force: bool = False, verbose: bool = True) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = _nvcc()
    hdrs = _headers()
    jobs = []
    for src in SOURCES:
        obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        if force or _stale(obj, [os.path.join(CSRC, src)] + hdrs):
            jobs.append([nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        return subprocess.run(cmd, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as pool:
        for res in pool.map(run, jobs):
            if verbose and (res.stdout or res.stderr):
                print(res.stdout + res.stderr, flush=True)
            if res.returncode != 0:
                raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-o", LIB] + \
           [os.path.join(OBJDIR, s.replace(".cu", ".o")) for s in SOURCES]
    res = run(link)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    print(LIB)
