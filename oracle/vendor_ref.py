"""Recipe that places the UNMODIFIED reference package under oracle/_ref/ (test / baseline infrastructure).

TEST INFRASTRUCTURE ONLY: nothing under gpmp_b200/ may import this module or oracle/_ref (same rule as the rest
of oracle/).  Only tests/, __graft_entry__.smoke(), bench.py's `cpu_baseline` leg and `bench.py --impl reference`
use it, and only as the checker / the CPU arm.

GPmp 0.9.37 is pure Python (pyproject.toml:1-3, no native sources), so "building" the reference is a copy of its
`gpmp/` package directory from where it lies (/root/reference) into oracle/_ref/gpmp.  oracle/_ref/ is listed in
.gitignore (reference sources never enter the history) but not in .gpurunignore, so the copy travels to the GPU
box, where /root/reference does not exist.  `python -m oracle.vendor_ref` (or __graft_entry__.build()) runs it.

    from oracle import vendor_ref
    gp = vendor_ref.import_reference()          # `import gpmp` from oracle/_ref with GPMP_BACKEND=torch
"""
from __future__ import annotations

import importlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("GPMP_REFERENCE_SRC", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DST, "gpmp", "__init__.py"))


def vendor(force: bool = False) -> bool:
    """Copy REF_SRC/gpmp -> oracle/_ref/gpmp (plus VERSION / LICENSE next to it).  Returns True when the copy
    exists afterwards.  A no-op on the GPU box (no REF_SRC there): the prebuilt copy is used as it came."""
    src = os.path.join(REF_SRC, "gpmp")
    if not os.path.isdir(src):
        return available()
    dst = os.path.join(REF_DST, "gpmp")
    if available() and not force:
        # refresh only when the source is newer than the copy
        newest = max(os.path.getmtime(os.path.join(r, f)) for r, _, fs in os.walk(src) for f in fs if f.endswith(".py"))
        if os.path.getmtime(os.path.join(dst, "__init__.py")) >= newest:
            return True
    os.makedirs(REF_DST, exist_ok=True)
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for extra in ("VERSION", "LICENSE", "LICENSE.txt", "COPYING"):
        f = os.path.join(REF_SRC, extra)
        if os.path.isfile(f):
            shutil.copy2(f, os.path.join(REF_DST, extra))
    os.utime(os.path.join(dst, "__init__.py"))
    return True


def ensure_plot_stub():
    """`import gpmp.mcmc` needs matplotlib to be importable (gpmp/mcmc/mh.py:53); when it is not installed, an
    import-only stand-in (oracle/stubs) goes at the END of sys.path."""
    import importlib.util

    if importlib.util.find_spec("matplotlib") is None:
        stubs = os.path.join(HERE, "stubs")
        if stubs not in sys.path:
            sys.path.append(stubs)


def import_reference(backend: str = "torch"):
    """Import the vendored reference as the top-level package `gpmp` with the given numerical backend.
    The backend is fixed at first import (gpmp/num/__init__.py:21-40): one backend per process."""
    if "gpmp" in sys.modules:
        mod = sys.modules["gpmp"]
        have = getattr(getattr(mod, "config", None), "get_config", lambda: None)()
        cur = getattr(have, "backend", backend)
        if cur != backend:
            raise RuntimeError(f"the reference is already imported with backend {cur!r}")
        return mod
    if not available() and not vendor():
        raise ImportError("oracle/_ref/gpmp is missing and /root/reference is not present: run "
                          "`python -m oracle.vendor_ref` where the reference tree exists")
    ensure_plot_stub()
    os.environ["GPMP_BACKEND"] = backend
    os.environ.setdefault("GPMP_LOG_LEVEL", "WARNING")
    if REF_DST not in sys.path:
        sys.path.insert(0, REF_DST)
    return importlib.import_module("gpmp")


if __name__ == "__main__":
    ok = vendor(force="--force" in sys.argv)
    print("oracle/_ref/gpmp", "ready" if ok else "NOT available (no reference tree here)")
