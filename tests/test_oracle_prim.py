"""The scipy/numpy primitives the oracle calls, pinned against the plain-C
restatement of their published algorithms (oracle/prim.c)."""

import ctypes
import os
import subprocess

import numpy as np
import pytest
import scipy.signal as sps

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_dp = ctypes.POINTER(ctypes.c_double)
_l = ctypes.c_long


def _p(a):
    return a.ctypes.data_as(_dp)


@pytest.fixture(scope="module")
def prim():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "liboracle_prim.so"))
    lib.prim_upfirdn_len.restype = _l
    lib.prim_upfirdn_len.argtypes = [_l, _l, _l, _l]
    lib.prim_convolve_full.argtypes = [_dp, _l, _dp, _l, _dp]
    lib.prim_sosfilt.argtypes = [_dp, _l, _dp, _l, _dp, _dp]
    lib.prim_lfilter.argtypes = [_dp, _l, _dp, _l, _dp, _l, _dp, _dp]
    lib.prim_upfirdn.argtypes = [_dp, _l, _dp, _l, _l, _l, _dp]
    lib.prim_rdft.argtypes = [_dp, _l, _dp, _dp]
    return lib


def test_convolution_definition(prim):
    rng = np.random.default_rng(0)
    x, h = rng.standard_normal(3000), rng.standard_normal(113)
    y = np.empty(len(x) + len(h) - 1)
    prim.prim_convolve_full(_p(x), len(x), _p(h), len(h), _p(y))
    assert np.max(np.abs(y - np.convolve(x, h, "full"))) < 1e-12 * np.max(np.abs(y))
    assert np.max(np.abs(y - sps.oaconvolve(x, h, "full"))) < 1e-12 * np.max(np.abs(y))


def test_sosfilt_is_df2t(prim):
    rng = np.random.default_rng(1)
    sos = np.ascontiguousarray(sps.butter(8, [1, 100], btype="bandpass", fs=5000, output="sos"))
    x = rng.standard_normal(20000) + 1.0
    zi = rng.standard_normal((sos.shape[0], 2))
    z = zi.copy()
    y = np.empty_like(x)
    prim.prim_sosfilt(_p(sos), sos.shape[0], _p(x), len(x), _p(z), _p(y))
    ry, rz = sps.sosfilt(sos, x, zi=zi)
    # same recurrence, same operation order: scipy matches bit for bit
    assert np.array_equal(y, ry) and np.array_equal(z, rz)


def test_lfilter_is_df2t(prim):
    rng = np.random.default_rng(2)
    b, a = sps.iirnotch(60, 10, fs=5000)
    b, a = np.ascontiguousarray(b), np.ascontiguousarray(a)
    x = rng.standard_normal(20000)
    zi = rng.standard_normal(2)
    z = zi.copy()
    y = np.empty_like(x)
    prim.prim_lfilter(_p(b), len(b), _p(a), len(a), _p(x), len(x), _p(z), _p(y))
    ry, rz = sps.lfilter(b, a, x, zi=zi)
    assert np.max(np.abs(y - ry)) < 1e-12 * np.max(np.abs(ry))
    assert np.max(np.abs(z - rz)) < 1e-12 * max(np.max(np.abs(rz)), 1.0)


@pytest.mark.parametrize("up,down", [(1, 20), (3, 7), (5, 1), (2, 3)])
def test_upfirdn_definition_and_length(prim, up, down):
    rng = np.random.default_rng(3)
    h = sps.firwin(61, 1.0 / max(up, down))
    x = rng.standard_normal(4001)
    ny = prim.prim_upfirdn_len(len(x), len(h), up, down)
    y = np.empty(ny)
    prim.prim_upfirdn(_p(h), len(h), _p(x), len(x), up, down, _p(y))
    ref = sps.upfirdn(h, x, up, down)
    assert ref.shape == y.shape                      # output length: bit exact
    assert np.max(np.abs(y - ref)) < 1e-12 * np.max(np.abs(ref))


@pytest.mark.parametrize("n", [8, 250, 1001, 1024])
def test_rfft_definition(prim, n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n)
    re, im = np.empty(n // 2 + 1), np.empty(n // 2 + 1)
    prim.prim_rdft(_p(x), n, _p(re), _p(im))
    ref = np.fft.rfft(x)
    assert np.max(np.abs(re + 1j * im - ref)) < 1e-11 * np.max(np.abs(ref))
