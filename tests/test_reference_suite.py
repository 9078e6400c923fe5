"""The reference's OWN acceptance tests for the hot path, run unmodified against
this package: its test
files are collected from the reference install with ``openseize.core.numerical``
and ``openseize.core.producer`` aliased to ``openseize_b200``'s, so the
reference's unmodified operator layer drives the CUDA kernels.

reference tests/test_oaconvolve.py:14-94, tests/test_iir.py:77-158,216-326,
tests/test_resampling.py:39-139, tests/test_spectra.py:16-615,
tests/test_pipelines.py, tests/test_concurrency.py:63-167 (picklability; the one
test that needs the downloadable demo recording is deselected: no network).
"""

import os
import subprocess
import sys

import pytest

from oracle import refload
from tests.conftest import ROOT, has_cuda

FILES = ["test_oaconvolve.py", "test_iir.py", "test_resampling.py", "test_spectra.py",
         "test_pipelines.py", "test_concurrency.py"]


def _run(fake, files, timeout):
    tdir = refload.tests_dir()
    if tdir is None or refload.location() is None:
        pytest.skip("reference not installed (baseline/_ref) -- run __graft_entry__.build() "
                    "where /root/reference exists")
    env = dict(os.environ, PYTHONPATH=ROOT, OSZ_REF_FAKE="1" if fake else "0")
    cmd = [sys.executable, "-m", "pytest", "-q", "-x", "-p", "tests.ref_alias_plugin",
           "-p", "no:cacheprovider", "--rootdir", tdir, "-k", "not test_edfreader"]
    cmd += [os.path.join(tdir, f) for f in files]
    res = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    tail = (res.stdout or "")[-3000:] + (res.stderr or "")[-2000:]
    assert res.returncode == 0, tail
    assert " passed" in res.stdout and "failed" not in res.stdout, tail
    return res.stdout


@pytest.mark.gpu
def test_reference_suite_on_gpu():
    assert has_cuda()
    out = _run(False, FILES, 1500)
    print(out[-400:])


def test_reference_suite_host_logic():
    """Same route on the CPU with the numpy stand-in kernels: proves the alias
    mechanism and the host logic; the arithmetic is checked on the GPU."""
    _run(True, ["test_oaconvolve.py", "test_resampling.py", "test_pipelines.py"], 900)
