"""API conformance: the reference's OWN unittest files (copied unmodified to oracle/_ref/tests by
oracle/build_ref.py) run against tempest_b200 on the GPU.

    python tools/ref_conformance.py [files...]      # default: the sampler-level files of SURVEY section 4

`import tempest` / `from tempest.sampler import Sampler` / `from tempest.config import SamplerConfig` resolve to
tempest_b200 through a sys.modules shim, so the tests construct tempest_b200.Sampler with their own plain-numpy
callables (the arbitrary-callable path: split propose / accept kernels around the user's functions) and check the
reference's contracts: returned keys / shapes / types, counters, posterior normalisation, evidence accuracy,
boundary conditions, resume.  Writes one line per test and a per-file summary to gpurun_out/ref_conformance.txt.
TEST INFRASTRUCTURE ONLY."""
import io
import os
import sys
import time
import unittest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tempest_b200  # noqa: E402
import tempest_b200.config  # noqa: E402
import tempest_b200.sampler  # noqa: E402

sys.modules["tempest"] = tempest_b200
sys.modules["tempest.sampler"] = tempest_b200.sampler
sys.modules["tempest.config"] = tempest_b200.config

TESTS = os.path.join(ROOT, "oracle", "_ref", "tests")
DEFAULT = ["test_sampler.py", "test_sample_method.py", "test_posterior_evidence.py", "test_end_to_end.py",
           "test_edge_cases.py", "test_sampler_features.py", "test_volume_variation.py", "test_state.py"]


def main():
    files = sys.argv[1:] or DEFAULT
    if not os.path.isdir(TESTS):
        print("oracle/_ref/tests is missing: run `python oracle/build_ref.py` where /root/reference exists")
        return 2
    sys.path.insert(0, TESTS)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    cwd = os.path.join(ROOT, "gpurun_out", "_conformance_cwd")      # the tests write states/ and *.state files
    os.makedirs(cwd, exist_ok=True)
    os.chdir(cwd)
    lines, summary = [], []
    for fname in files:
        mod = fname[:-3]
        t0 = time.time()
        try:
            suite = unittest.defaultTestLoader.loadTestsFromName(mod)
        except Exception as exc:       # import error inside the test module
            summary.append(f"{fname}: IMPORT ERROR {type(exc).__name__}: {exc}")
            continue
        res = unittest.TextTestRunner(stream=io.StringIO(), verbosity=0).run(suite)
        bad = {t.id(): ("FAIL", tb) for t, tb in res.failures}
        bad.update({t.id(): ("ERROR", tb) for t, tb in res.errors})
        for tid, (kind, tb) in sorted(bad.items()):
            last = [ln for ln in tb.strip().splitlines() if ln.strip()][-1]
            lines.append(f"{kind} {tid}: {last[:300]}")
        summary.append(f"{fname}: ran {res.testsRun}, failed {len(res.failures)}, errors {len(res.errors)}, "
                       f"skipped {len(res.skipped)} in {time.time() - t0:.1f} s")
    out = "\n".join(["# reference unittest files run against tempest_b200 (tools/ref_conformance.py)"] + summary +
                    ["", "# failing tests"] + (lines or ["(none)"])) + "\n"
    with open(os.path.join(ROOT, "gpurun_out", "ref_conformance.txt"), "w") as f:
        f.write(out)
    print(out)
    return 0


if __name__ == "__main__":
    sys.exit(main())
