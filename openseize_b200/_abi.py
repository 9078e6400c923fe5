"""ctypes binding of ``include/osz_b200.h`` (the C ABI of the CUDA library).

There is no CPU implementation behind this module: if the shared library has
not been built, or no CUDA device is visible, the operators raise.
"""

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libosz_b200.so")

OSZ_OK, OSZ_ERR_ARG, OSZ_ERR_CUDA, OSZ_ERR_UNSUPPORTED, OSZ_ERR_ALLOC = 0, -1, -2, -3, -4
FIR_AUTO, FIR_DIRECT, FIR_FFT, FIR_FFT_F32 = 0, 1, 2, 3
DETREND = {None: 0, False: 0, "none": 0, "constant": 1, "linear": 2}

_i64, _dp, _vp = c_int64, POINTER(c_double), c_void_p

# name -> (restype, argtypes); every symbol declared in include/osz_b200.h
SIGNATURES = {
    "osz_version": (c_int, []),
    "osz_last_error": (c_char_p, []),
    "osz_device_info": (c_int, [c_int, POINTER(c_int), POINTER(c_int), POINTER(c_int),
                                POINTER(_i64), POINTER(_i64)]),
    "osz_launch_count": (_i64, []),
    "osz_dev_malloc": (c_int, [POINTER(_vp), c_size_t]),
    "osz_dev_free": (c_int, [_vp]),
    "osz_host_alloc": (c_int, [POINTER(_vp), c_size_t]),
    "osz_host_free": (c_int, [_vp]),
    "osz_stream_create": (c_int, [POINTER(_vp)]),
    "osz_stream_destroy": (c_int, [_vp]),
    "osz_stream_sync": (c_int, [_vp]),
    "osz_memcpy_h2d_async": (c_int, [_vp, _vp, c_size_t, _vp]),
    "osz_memcpy_d2h_async": (c_int, [_vp, _vp, c_size_t, _vp]),
    "osz_memcpy_d2d_async": (c_int, [_vp, _vp, c_size_t, _vp]),
    "osz_memcpy2d_h2d_async": (c_int, [_vp, c_size_t, _vp, c_size_t, c_size_t, c_size_t, _vp]),
    "osz_memcpy2d_d2h_async": (c_int, [_vp, c_size_t, _vp, c_size_t, c_size_t, c_size_t, _vp]),
    "osz_memset_async": (c_int, [_vp, c_int, c_size_t, _vp]),
    "osz_pack_rows_f64": (c_int, [_vp, _i64, _i64, _i64, _vp, _i64, _vp]),
    "osz_unpack_rows_f64": (c_int, [_vp, _i64, _i64, _i64, _i64, _vp, _vp]),
    "osz_unpack_rows_c128": (c_int, [_vp, _i64, _i64, _i64, _i64, _vp, _vp]),
    "osz_zip_complex_f64": (c_int, [_vp, _i64, _vp, _i64, _i64, _i64, _vp, _vp]),
    "osz_widen_f32_f64": (c_int, [_vp, _vp, _i64, _vp]),
    "osz_widen_i16_f64": (c_int, [_vp, _vp, _i64, _vp]),
    "osz_widen_rows_f32_f64": (c_int, [_vp, _i64, _vp, _i64, _i64, _i64, _vp]),
    "osz_widen_rows_i16_f64": (c_int, [_vp, _i64, _vp, _i64, _i64, _i64, _vp]),
    "osz_decode_edf_records_f64": (c_int, [_vp, _i64, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _i64,
                                           _i64, _vp]),
    "osz_fir_plan_create": (c_int, [POINTER(_vp), _dp, c_int, c_int]),
    "osz_fir_plan_destroy": (c_int, [_vp]),
    "osz_fir_plan_algo": (c_int, [_vp]),
    "osz_fir_exec_f64": (c_int, [_vp, _vp, _i64, _i64, _i64, _vp, _i64, _vp]),
    "osz_sos_plan_create": (c_int, [POINTER(_vp), _dp, c_int]),
    "osz_sos_plan_destroy": (c_int, [_vp]),
    "osz_sos_exec_f64": (c_int, [_vp, _vp, _i64, _i64, _i64, c_int, _vp, _vp, _i64, _vp]),
    "osz_sos_state_from_sample_f64": (c_int, [_vp, _dp, _vp, _i64, _i64, _i64, _vp, _vp]),
    "osz_tf_plan_create": (c_int, [POINTER(_vp), _dp, c_int, _dp, c_int]),
    "osz_tf_plan_destroy": (c_int, [_vp]),
    "osz_tf_plan_states": (c_int, [_vp]),
    "osz_tf_exec_f64": (c_int, [_vp, _vp, _i64, _i64, _i64, c_int, _vp, _vp, _i64, _vp]),
    "osz_tf_state_from_sample_f64": (c_int, [_vp, _dp, _vp, _i64, _i64, _i64, _vp, _vp]),
    "osz_upfirdn_plan_create": (c_int, [POINTER(_vp), _dp, c_int, c_int, c_int]),
    "osz_upfirdn_plan_destroy": (c_int, [_vp]),
    "osz_upfirdn_exec_f64": (c_int, [_vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _i64,
                                     _vp]),
    "osz_upfirdn_plan_set_compute": (c_int, [_vp, c_int]),
    "osz_upfirdn_plan_compute": (c_int, [_vp]),
    "osz_fir_exec_f32": (c_int, [_vp, _vp, _i64, _i64, _i64, _vp, _i64, _vp]),
    "osz_sos_exec_f32": (c_int, [_vp, _vp, _i64, _i64, _i64, c_int, _vp, _vp, _i64, _vp]),
    "osz_sos_state_from_sample_f32": (c_int, [_vp, _dp, _vp, _i64, _i64, _i64, _vp, _vp]),
    "osz_sos_lookahead_f64": (c_int, [_vp, _dp, _vp, _i64, _i64, _i64, c_int, _vp, _vp]),
    "osz_sos_lookahead_f32": (c_int, [_vp, _dp, _vp, _i64, _i64, _i64, c_int, _vp, _vp]),
    "osz_upfirdn_exec_f32": (c_int, [_vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _i64,
                                     _vp]),
    "osz_welch_accum_f32": (c_int, [_vp, _vp, _i64, _i64, _i64, _vp, _i64, _vp]),
    "osz_narrow_rows_f64_f32": (c_int, [_vp, _i64, _vp, _i64, _i64, _i64, _vp]),
    "osz_widen_rows_i16_f32": (c_int, [_vp, _i64, _vp, _i64, _i64, _i64, _vp]),
    "osz_sos_tail_state_f64": (c_int, [_vp, _vp, _i64, _i64, _i64, c_int, _vp, _vp]),
    "osz_sos_plan_settle": (_i64, [_vp]),
    "osz_sos_plan_has_weights": (c_int, [_vp]),
    "osz_sosdec_spans": (c_int, [_vp, _vp, _i64, _i64]),
    "osz_sosdec_exec_f64": (c_int, [_vp, _vp, _vp, _i64, _i64, _i64, c_int, _vp, c_int, _i64, _vp,
                                    _i64, _i64, _i64, _vp, _vp]),
    "osz_sosdec_boundary_f64": (c_int, [_vp, _vp, c_int, c_int, _vp, _i64, c_int, _i64, _i64, _i64,
                                        _vp, _i64, _i64, _i64, _i64, _i64, _vp]),
    "osz_upfirdn_plan_set_kernel": (c_int, [_vp, c_int]),
    "osz_upfirdn_plan_kernel": (c_int, [_vp]),
    "osz_spec_plan_create": (c_int, [POINTER(_vp), c_int, c_int, _dp, c_int, c_double]),
    "osz_spec_plan_destroy": (c_int, [_vp]),
    "osz_spec_plan_path": (c_int, [_vp]),
    "osz_spec_plan_set_compute": (c_int, [_vp, c_int, _dp]),
    "osz_spec_plan_compute": (c_int, [_vp]),
    "osz_welch_accum_f64": (c_int, [_vp, _vp, _i64, _i64, _i64, _vp, _i64, _vp]),
    "osz_periodogram_f64": (c_int, [_vp, _vp, _i64, _i64, _i64, _vp, _vp]),
    "osz_stft_f64": (c_int, [_vp, _vp, _i64, _i64, _i64, _vp, _vp]),
    "osz_spec_prepare_f64": (c_int, [_vp, _i64, _i64, _i64, _i64, _vp, c_int, _vp, _i64, _vp]),
    "osz_take_cols_f64": (c_int, [_vp, _i64, _i64, _vp, _i64, _vp, _i64, _vp]),
    "osz_row_moments_slots": (c_int, []),
    "osz_row_moments_f64": (c_int, [_vp, _i64, _i64, _i64, c_int, _vp, _vp, _vp]),
    "osz_row_standardize_f64": (c_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _i64, _vp]),
    "osz_col_moments_f64": (c_int, [_vp, _i64, _i64, _i64, c_int, _vp, _vp, _vp, _i64, _vp]),
}

_lib = None


def load():
    """Load libosz_b200.so (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            # Not a fallback: the only thing tried is compiling the CUDA library
            # itself (nvcc, sm_100a).  Without nvcc the operators cannot run.
            try:
                from openseize_b200.csrc import build as _build

                _build.build()
            except Exception as exc:
                raise RuntimeError(
                    "openseize_b200: %s is missing and could not be built (%s) -- build it "
                    "with `python -m openseize_b200.csrc.build` (or __graft_entry__.build()). "
                    "There is no CPU fallback." % (LIB_PATH, exc))
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error():
    return load().osz_last_error().decode("utf-8", "replace")


def check(rc, what=""):
    """Translate a status code into the Python exception the reference-facing
    API documents."""
    if rc == OSZ_OK:
        return
    msg = "%s%s" % (what + ": " if what else "", last_error())
    if rc == OSZ_ERR_ARG:
        raise ValueError(msg)
    if rc == OSZ_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)


def launch_count():
    return int(load().osz_launch_count())


def as_double_array(values):
    """Contiguous C double array (host) from a sequence / ndarray."""
    import numpy as np

    arr = np.ascontiguousarray(values, dtype=np.float64)
    return arr, arr.ctypes.data_as(_dp)
