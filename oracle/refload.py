"""Locate and import the UNMODIFIED reference (mscaudill/openseize).

TEST INFRASTRUCTURE ONLY -- used by ``bench.py --impl reference`` / its
``cpu_baseline`` leg, ``tests/test_reference_suite.py`` and
``oracle/make_golden.py``; nothing under ``openseize_b200/`` imports this.

The reference is pure Python.  ``install()`` (run by ``__graft_entry__.build()``
in the build container, where ``/root/reference`` exists) pip-installs it into
the git-ignored ``baseline/_ref`` so that it travels to the GPU box with the
gpurun snapshot, and puts a copy of the reference's own test files next to it
(``baseline/_ref/_tests``) for the acceptance run against this package.  Its
filtering modules import matplotlib for their plotting mixins
(``filtering/mixins.py:11,14``); matplotlib is absent from this image, so empty
stand-in modules are registered when it cannot be imported.
"""

import importlib
import os
import shutil
import subprocess
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
INSTALLED = os.path.join(ROOT, "baseline", "_ref")
CHECKOUT = "/root/reference"


def location():
    """Directory to put on sys.path for ``import openseize``, or None."""
    if os.path.isdir(os.path.join(INSTALLED, "openseize")):
        return INSTALLED
    if os.path.isdir(os.path.join(CHECKOUT, "src", "openseize")):
        return os.path.join(CHECKOUT, "src")
    return None


def tests_dir():
    for cand in (os.path.join(INSTALLED, "_tests"), os.path.join(CHECKOUT, "tests")):
        if os.path.isdir(cand):
            return cand
    return None


def stub_plotting():
    """Stand-ins for the plotting imports of the reference when matplotlib is
    not installed (only names the reference touches at import time)."""
    try:
        importlib.import_module("matplotlib.pyplot")
        return False
    except Exception:
        pass
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.widgets"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib.pyplot"].Axes = object
    sys.modules["matplotlib.patches"].Rectangle = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].patches = sys.modules["matplotlib.patches"]
    sys.modules["matplotlib"].widgets = sys.modules["matplotlib.widgets"]
    return True


def load():
    """Import the reference; returns the ``openseize`` package or None."""
    where = location()
    if where is None:
        return None
    stub_plotting()
    if where not in sys.path:
        sys.path.insert(0, where)
    import openseize  # noqa: F401  (the reference, not openseize_b200)
    import openseize.filtering.fir  # noqa: F401
    import openseize.filtering.iir  # noqa: F401
    import openseize.resampling.resampling  # noqa: F401
    import openseize.spectra.estimators  # noqa: F401
    return sys.modules["openseize"]


def install(force=False):
    """pip-install the read-only checkout into baseline/_ref (from a scratch copy:
    the build writes egg-info into the source tree) and copy its test files.
    No-op without /root/reference.  Returns the install directory or None."""
    if not os.path.isdir(CHECKOUT):
        return INSTALLED if os.path.isdir(os.path.join(INSTALLED, "openseize")) else None
    if os.path.isdir(os.path.join(INSTALLED, "openseize")) and not force:
        if os.path.isdir(os.path.join(INSTALLED, "_tests")):
            return INSTALLED
    scratch = "/tmp/osz_ref_copy"
    shutil.rmtree(scratch, ignore_errors=True)
    shutil.copytree(CHECKOUT, scratch, ignore=shutil.ignore_patterns(".git"))
    os.makedirs(os.path.dirname(INSTALLED), exist_ok=True)
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation",
           "--no-deps", "--find-links", "/opt/wheelhouse", "--upgrade", "--target", INSTALLED,
           scratch]
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    tdst = os.path.join(INSTALLED, "_tests")
    shutil.rmtree(tdst, ignore_errors=True)
    shutil.copytree(os.path.join(CHECKOUT, "tests"), tdst,
                    ignore=shutil.ignore_patterns("__pycache__"))
    shutil.rmtree(scratch, ignore_errors=True)
    return INSTALLED


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
