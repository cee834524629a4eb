"""Recipe for ``oracle/_ref/``: an unmodified copy of the reference's ``tempest`` package.

    python oracle/build_ref.py [--ref /root/reference]

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  The reference is pure Python: "building" it is copying
``/root/reference/tempest`` (its 15 source files, byte for byte) to ``oracle/_ref/tempest`` so that it
travels to the GPU box, where ``/root/reference`` does not exist.  ``oracle/_ref/`` is git-ignored (the
reference's sources never enter this repository's history) but not gpurun-ignored.  Only ``bench.py``'s
``--impl reference`` / ``cpu_baseline`` legs and the tests import from it; nothing under ``tempest_b200/``
does (``tests/test_host_logic.py`` enforces that).  ``__graft_entry__.build()`` runs this whenever
``/root/reference`` is present; ``MANIFEST.json`` records the sha256 of every copied file next to the sha256
of its source so that "unmodified" can be checked.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "oracle", "_ref")


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def build(ref_root: str = "/root/reference", verbose: bool = True) -> bool:
    """Copy the package; returns False (and leaves any existing copy alone) when the reference is absent."""
    src = os.path.join(ref_root, "tempest")
    if not os.path.isdir(src):
        return False
    dst = os.path.join(DEST, "tempest")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(DEST, exist_ok=True)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    manifest = {}
    for base, _, files in os.walk(dst):
        for name in sorted(files):
            p = os.path.join(base, name)
            rel = os.path.relpath(p, dst)
            manifest[rel] = {"sha256": _sha(p), "source_sha256": _sha(os.path.join(src, rel))}
            assert manifest[rel]["sha256"] == manifest[rel]["source_sha256"], rel
    # the reference's own unittest files ride along (tools/ref_conformance.py runs them against tempest_b200.Sampler)
    tsrc, tdst = os.path.join(ref_root, "tests"), os.path.join(DEST, "tests")
    if os.path.isdir(tsrc):
        if os.path.isdir(tdst):
            shutil.rmtree(tdst)
        shutil.copytree(tsrc, tdst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "files": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print(f"oracle/_ref: copied {len(manifest)} files of the unmodified reference from {src}")
    return True


def import_reference():
    """``import tempest`` from oracle/_ref (raises ImportError when the copy is missing)."""
    if not os.path.isdir(os.path.join(DEST, "tempest")):
        raise ImportError("oracle/_ref/tempest is missing: run `python oracle/build_ref.py` where /root/reference exists")
    if DEST not in sys.path:
        sys.path.insert(0, DEST)
    import tempest  # noqa: F401

    if not os.path.abspath(tempest.__file__).startswith(DEST):
        raise ImportError(f"`tempest` resolved to {tempest.__file__}, not to oracle/_ref")
    return tempest


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ok = build(ap.parse_args().ref)
    sys.exit(0 if ok else 1)
