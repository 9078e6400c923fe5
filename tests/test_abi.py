"""The C-ABI library loads and exports every symbol include/osz_b200.h
declares (no compute calls: there is no GPU here)."""

import ctypes
import os
import re

import pytest

from openseize_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "osz_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(osz_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_abi.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    return ctypes.CDLL(_abi.LIB_PATH)


def test_header_and_binding_agree():
    assert _declared() == sorted(_abi.SIGNATURES), "include/osz_b200.h and _abi.py differ"


def test_every_declared_symbol_is_exported(lib):
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, missing


def test_version_and_error_string(lib):
    lib.osz_version.restype = ctypes.c_int
    assert lib.osz_version() >= 100
    lib.osz_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.osz_last_error(), bytes)
    lib.osz_launch_count.restype = ctypes.c_int64
    assert lib.osz_launch_count() == 0


def test_bad_arguments_are_rejected_without_a_gpu(lib):
    lib.osz_fir_plan_create.restype = ctypes.c_int
    handle = ctypes.c_void_p()
    rc = lib.osz_fir_plan_create(ctypes.byref(handle), None, 0, 0)
    assert rc == _abi.OSZ_ERR_ARG and b"bad arguments" in lib.osz_last_error()
    rc = lib.osz_sos_plan_create(ctypes.byref(handle), None, 0)
    assert rc == _abi.OSZ_ERR_ARG
