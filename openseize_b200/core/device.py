"""Device plumbing for the hot path: plans, kernel launches, and the pinned
host <-> device chunk streaming under the producer iterator.

torch is used only for device memory, streams and events; every arithmetic
kernel is one of ours, called through the C ABI (``openseize_b200._abi``).
Signals live on the device as time-contiguous rows ``(rows, n)``: for a chunk
of shape ``(..., n, ...)`` with the sample axis at ``axis``, ``rows =
outer * inner`` and row ``o * inner + i`` holds ``chunk[o, :, i]``.
"""

import ctypes
import os
import math
from collections import OrderedDict

import numpy as np

from openseize_b200 import _abi

_torch = None
DEVICE = "cuda"


def torch():
    global _torch
    if _torch is None:
        import torch as _t

        _torch = _t
    return _torch


def require_cuda():
    """The hot path has no CPU implementation: fail loudly without a GPU."""
    t = torch()
    if not t.cuda.is_available():
        raise RuntimeError(
            "openseize_b200 runs its DSP kernels on a CUDA device (sm_100a) and "
            "none is visible; there is no CPU fallback.")
    _abi.load()
    return t


def _vp(addr):
    return ctypes.c_void_p(int(addr))


def _cur_stream():
    return _vp(torch().cuda.current_stream().cuda_stream)


def _rows_ptr(t):
    """(pointer, leading dimension) of a 2-D float64 device tensor whose rows
    are time-contiguous."""
    assert t.dim() == 2 and t.dtype in (torch().float64, torch().float32) and t.is_cuda
    if t.shape[1] > 1:
        assert t.stride(1) == 1, "rows must be time-contiguous"
    ld = t.stride(0) if t.shape[0] > 1 else max(t.shape[1], 1)
    return _vp(t.data_ptr()), int(ld)


# --------------------------------------------------------------------------
# side streams
# --------------------------------------------------------------------------
class _Streams:
    _by_device = {}

    @classmethod
    def get(cls):
        t = torch()
        dev = t.cuda.current_device()
        if dev not in cls._by_device:
            cls._by_device[dev] = (t.cuda.Stream(), t.cuda.Stream())
        return cls._by_device[dev]


_named_streams = {}


def side_stream(name):
    """A per-device CUDA stream for a stage that may run ahead of the stream its
    consumer launches on (e.g. the forward pass of a forward-backward filter); None
    off-GPU (test stand-ins)."""
    if DEVICE != "cuda":
        return None
    t = torch()
    key = (t.cuda.current_device(), name)
    if key not in _named_streams:
        _named_streams[key] = t.cuda.Stream()
    return _named_streams[key]


class on_stream:
    """``with on_stream(s):`` -- launches and allocations go to stream ``s`` (no-op for
    None).  Never held across a ``yield``."""

    def __init__(self, stream):
        self.ctx = torch().cuda.stream(stream) if stream is not None else None

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


# --------------------------------------------------------------------------
# layout of a chunk
# --------------------------------------------------------------------------
class Layout:
    """Where the sample axis sits in the reference-facing ndarray."""

    def __init__(self, shape, axis):
        shape = tuple(int(s) for s in shape)
        axis = int(np.arange(len(shape))[axis])
        self.axis = axis
        self.lead = shape[:axis]
        self.trail = shape[axis + 1:]
        self.outer = int(np.prod(self.lead, dtype=np.int64)) if self.lead else 1
        self.inner = int(np.prod(self.trail, dtype=np.int64)) if self.trail else 1
        self.rows = self.outer * self.inner

    def host_shape(self, n):
        return self.lead + (int(n),) + self.trail


_calib_cache = None      # _Cache(64), made below once the class exists
_copy_pool = None


def _staged_copy(dst, src):
    """Copy a host chunk into its pinned staging block.  One memcpy thread moves
    ~10 GB/s, a fifth of what the PCIe link takes (measured: 1.0 G channel-samples/s
    of pageable float64 into psd); large chunks are split over a few threads
    (numpy releases the GIL while it copies)."""
    global _copy_pool
    if src.nbytes < (8 << 20) or src.ndim != 2 or src.shape[1] < 4096:
        np.copyto(dst, src)
        return
    import concurrent.futures as cf

    if _copy_pool is None:
        _copy_pool = cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1),
                                           thread_name_prefix="osz-stage")
    parts = _copy_pool._max_workers
    n = src.shape[1]
    step = -(-n // parts)
    futs = [_copy_pool.submit(np.copyto, dst[:, a:a + step], src[:, a:a + step])
            for a in range(0, n, step)]
    for f in futs:
        f.result()


def _calibration(chunk):
    """Device copies of an EDF chunk's (channel offsets, slopes, offsets), cached
    per value (they are the same for every chunk of a reader)."""
    t = torch()
    key = (chunk.chan_off.tobytes(), chunk.slopes.tobytes(), chunk.offsets.tobytes())
    # keyed by device like the plans, least recently used entry evicted (tensors still
    # referenced by queued kernels are kept alive by torch's stream-ordered allocator)
    return _calib_cache.get(key, lambda: (t.from_numpy(chunk.chan_off.copy()).to("cuda"),
                                          from_host(chunk.slopes), from_host(chunk.offsets)))


def _upload_edf(chunk, layout, alloc=None):
    """EDF records -> device rows: the int16 records cross PCIe as they are (one
    contiguous copy from pinned memory on the side stream), a kernel
    de-interleaves and calibrates them into float64 rows."""
    t = require_cuda()
    rows, n = chunk.shape
    dev = alloc(rows, n) if alloc is not None else t.empty((rows, n), dtype=t.float64,
                                                           device="cuda")
    if n == 0 or rows == 0:
        return dev
    h2d, _ = _Streams.get()
    cur = t.cuda.current_stream()
    off, sl, of = _calibration(chunk)
    with t.cuda.stream(h2d):
        raw = t.empty(chunk.records.shape, dtype=t.int16, device="cuda")
        src = t.from_numpy(chunk.records)
        if not src.is_pinned():
            stage = t.empty(chunk.records.shape, dtype=t.int16, pin_memory=True)
            stage.copy_(src)
            src = stage
        raw.copy_(src, non_blocking=True)
        raw._osz_keepalive = src
    cur.wait_stream(h2d)
    raw.record_stream(cur)
    ldd = dev.stride(0) if rows > 1 else n
    _abi.check(_abi.load().osz_decode_edf_records_f64(
        _vp(raw.data_ptr()), chunk.records.shape[1], chunk.spr, _vp(off.data_ptr()),
        _vp(sl.data_ptr()), _vp(of.data_ptr()), chunk.skip, _vp(dev.data_ptr()), ldd, rows, n,
        _cur_stream()), "decode_edf")
    return dev


def upload(arr, layout, alloc=None):
    """Host chunk -> device rows ``(rows, n)`` of the I/O type on the current stream
    (float64; float32 in the float32 I/O mode, where float64 / int16 host chunks are
    converted to float32 -- float64 on the host, int16 on the device)."""
    if IO != "float32":
        return _upload_f64(arr, layout, alloc)
    t = require_cuda()
    raw_chunk = hasattr(arr, "records") and hasattr(arr, "chan_off")
    a = None if raw_chunk else np.asarray(arr)
    if raw_chunk or layout.inner != 1 or a.dtype not in (np.float32, np.float64, np.int16):
        # EDF records / sample axis not last / other dtypes: the float64 path, then narrow
        dev64 = _upload_f64(arr, layout, None)
        rows, n = dev64.shape
        return to_io(dev64, alloc(rows, n) if alloc is not None else None)
    n = a.shape[layout.axis]
    if a.dtype == np.float64:
        a = a.astype(np.float32)             # half the PCIe traffic
    a2 = np.ascontiguousarray(a.reshape(layout.outer, n))
    h2d, _ = _Streams.get()
    cur = t.cuda.current_stream()
    tdt = t.float32 if a2.dtype == np.float32 else t.int16
    src = t.from_numpy(a2) if a2.flags.writeable else t.from_numpy(a2.copy())
    if a2.dtype == np.float32:
        dev = alloc(layout.outer, n) if alloc is not None else None
        ev = getattr(alloc, "ready_event", None)
        if ev is not None:
            h2d.wait_event(ev)
        elif alloc is not None:
            h2d.wait_stream(cur)
        with t.cuda.stream(h2d):
            if dev is None:
                dev = t.empty((layout.outer, n), dtype=t.float32, device="cuda")
            if not src.is_pinned():
                stage = t.empty((layout.outer, n), dtype=tdt, pin_memory=True)
                stage.copy_(src)
                src = stage
            dev.copy_(src, non_blocking=True)
            dev._osz_keepalive = src
        cur.wait_stream(h2d)
        dev.record_stream(cur)
        return dev
    with t.cuda.stream(h2d):                 # int16: widened to float32 on the device
        raw = t.empty((layout.outer, n), dtype=tdt, device="cuda")
        if not src.is_pinned():
            stage = t.empty((layout.outer, n), dtype=tdt, pin_memory=True)
            stage.copy_(src)
            src = stage
        raw.copy_(src, non_blocking=True)
        raw._osz_keepalive = src
    cur.wait_stream(h2d)
    raw.record_stream(cur)
    dev = alloc(layout.outer, n) if alloc is not None else empty_rows((layout.outer, n))
    ldd = dev.stride(0) if layout.outer > 1 else n
    _abi.check(_abi.load().osz_widen_rows_i16_f32(_vp(raw.data_ptr()), n, _vp(dev.data_ptr()), ldd,
                                                  layout.outer, n, _cur_stream()), "widen")
    return dev


def _upload_f64(arr, layout, alloc=None):
    """Host chunk -> device rows ``(rows, n)`` float64 on the current stream.
    ``alloc(rows, n)`` places the rows in the consumer's staging ring.

    The copy is issued on a side stream from pinned memory (the chunk itself
    when it already is pinned, else a pinned staging block from torch's
    caching host allocator) so it overlaps the kernels of the previous chunk.
    """
    t = require_cuda()
    if hasattr(arr, "records") and hasattr(arr, "chan_off"):
        # an EDF RawChunk (file_io/edf.py): whole int16 records + calibration
        if layout.inner != 1 or layout.outer != arr.shape[0]:
            arr = arr.decode()
        else:
            return _upload_edf(arr, layout, alloc)
    a = np.asarray(arr)
    narrow = a.dtype in (np.float32, np.int16) and layout.inner == 1
    if a.dtype != np.float64 and not narrow:
        a = a.astype(np.float64)
    n = a.shape[layout.axis]
    h2d, _ = _Streams.get()
    cur = t.cuda.current_stream()
    if narrow:
        # float32 / int16 recordings (EDF samples are int16) cross PCIe in their
        # own width and are widened to float64 on the device -- the reference
        # returns float64 for every input dtype.
        tdt = t.float32 if a.dtype == np.float32 else t.int16
        a2 = np.ascontiguousarray(a.reshape(layout.outer, n))
        with t.cuda.stream(h2d):
            raw = t.empty((layout.outer, n), dtype=tdt, device="cuda")
            src = t.from_numpy(a2) if a2.flags.writeable else t.from_numpy(a2.copy())
            if not src.is_pinned():
                stage = t.empty((layout.outer, n), dtype=tdt, pin_memory=True)
                stage.copy_(src)
                src = stage
            raw.copy_(src, non_blocking=True)
            raw._osz_keepalive = src
        cur.wait_stream(h2d)
        raw.record_stream(cur)
        dev = alloc(layout.outer, n) if alloc is not None else t.empty(
            (layout.outer, n), dtype=t.float64, device="cuda")
        lib = _abi.load()
        ldd = dev.stride(0) if layout.outer > 1 else n
        fn = lib.osz_widen_rows_f32_f64 if a.dtype == np.float32 else lib.osz_widen_rows_i16_f64
        _abi.check(fn(_vp(raw.data_ptr()), n, _vp(dev.data_ptr()), ldd, layout.outer, n,
                      _cur_stream()), "widen")
        return dev

    def fence():
        # The H2D copy runs on a side stream so it overlaps the kernels of the
        # previous chunk.  It only has to wait for whoever used the destination
        # memory before: nothing for blocks allocated on the copy stream itself
        # (torch's allocator orders their reuse), the ring's allocation event for
        # space inside a consumer's staging ring.
        ev = getattr(alloc, "ready_event", None)
        if ev is not None:
            h2d.wait_event(ev)
        elif alloc is not None:
            h2d.wait_stream(cur)

    def fresh(shape):
        with t.cuda.stream(h2d):
            return t.empty(shape, dtype=t.float64, device="cuda")

    if layout.inner == 1:
        a2 = a.reshape(layout.outer, n)
        dev = alloc(layout.outer, n) if alloc is not None else fresh((layout.outer, n))
        ldd = dev.stride(0) if layout.outer > 1 else n
        fence()
        with t.cuda.stream(h2d):
            direct = False
            if a2.strides[1] == 8 and a2.strides[0] >= n * 8 and a2.flags.writeable:
                try:
                    direct = t.from_numpy(a2).is_pinned()
                except Exception:      # exotic strides torch refuses to wrap
                    direct = False
            if direct:
                rc = _abi.load().osz_memcpy2d_h2d_async(
                    _vp(dev.data_ptr()), ldd * 8, _vp(a2.ctypes.data), a2.strides[0], n * 8,
                    layout.outer, _vp(h2d.cuda_stream))
                _abi.check(rc, "upload")
                dev._osz_keepalive = a2
            else:
                stage = t.empty((layout.outer, n), dtype=t.float64, pin_memory=True)
                _staged_copy(stage.numpy(), a2)
                dev.copy_(stage, non_blocking=True)
        cur.wait_stream(h2d)
        dev.record_stream(cur)
        return dev
    # sample axis is not last: ship the chunk as it is, transpose on the device
    flat = np.ascontiguousarray(a).reshape(layout.outer, n, layout.inner)
    stage = t.empty(flat.shape, dtype=t.float64, pin_memory=True)
    np.copyto(stage.numpy(), flat)
    raw = fresh(flat.shape)
    with t.cuda.stream(h2d):
        raw.copy_(stage, non_blocking=True)
    cur.wait_stream(h2d)
    raw.record_stream(cur)
    if alloc is not None:
        dev = alloc(layout.rows, n)
    else:
        dev = t.empty((layout.rows, n), dtype=t.float64, device="cuda")
    ldd = dev.stride(0) if layout.rows > 1 else n
    rc = _abi.load().osz_pack_rows_f64(_vp(raw.data_ptr()), layout.outer, n, layout.inner,
                                       _vp(dev.data_ptr()), ldd, _cur_stream())
    _abi.check(rc, "pack_rows")
    return dev


class Pending:
    """A device -> pinned-host copy in flight; ``get()`` waits and returns the
    ndarray in the reference's layout (backed by the pinned block)."""

    def __init__(self, host, event, shape, complex_=False, cast=None):
        self._host, self._event, self._shape, self._complex = host, event, shape, complex_
        self._cast = cast

    def get(self):
        self._event.synchronize()
        out = self._host.numpy()
        if self._complex:
            out = out.view(np.complex128)
        out = out.reshape(self._shape)
        return out.astype(self._cast) if self._cast is not None else out


class _PendingBlock:
    def __init__(self, host, event):
        self._host, self._event = host, event

    def get(self):
        self._event.synchronize()
        return self._host.numpy()


def download_block(dev):
    """A contiguous device tensor -> pinned host ndarray of the same shape, copied
    on the download stream; ``get()`` waits for it."""
    t = require_cuda()
    cur = t.cuda.current_stream()
    _, d2h = _Streams.get()
    dev = dev.contiguous()
    host = t.empty(tuple(dev.shape), dtype=dev.dtype, pin_memory=True)
    d2h.wait_stream(cur)
    with t.cuda.stream(d2h):
        host.copy_(dev, non_blocking=True)
        ev = t.cuda.Event()
        ev.record(d2h)
    dev.record_stream(d2h)
    return _PendingBlock(host, ev)


def download(dev, layout, complex_=False):
    """Device rows ``(rows, n)`` (float64, or complex128 stored as (rows, n, 2))
    -> :class:`Pending` host array of shape ``layout.host_shape(n)``."""
    t = require_cuda()
    n = dev.shape[1]
    cur = t.cuda.current_stream()
    _, d2h = _Streams.get()
    cast = None
    if IO == "float32":
        # float32 rows go to the host as they are when the sample axis is last; other
        # layouts and complex results take the float64 path and are narrowed on the host
        if dev.dtype == t.float32 and layout.inner == 1 and not complex_:
            host = t.empty(tuple(dev.shape), dtype=t.float32, pin_memory=True)
            d2h.wait_stream(cur)
            with t.cuda.stream(d2h):
                host.copy_(dev, non_blocking=True)
                ev = t.cuda.Event()
                ev.record(d2h)
            dev.record_stream(d2h)
            return Pending(host, ev, layout.host_shape(n))
        cast = np.complex64 if complex_ else np.float32
        if dev.dtype == t.float32:
            dev = to_f64(dev)
    if layout.inner > 1:
        shape = (layout.outer, n, layout.inner) + ((2,) if complex_ else ())
        tmp = t.empty(shape, dtype=t.float64, device="cuda")
        src = dev.contiguous()
        fn = _abi.load().osz_unpack_rows_c128 if complex_ else _abi.load().osz_unpack_rows_f64
        rc = fn(_vp(src.data_ptr()), n, layout.outer, n, layout.inner, _vp(tmp.data_ptr()),
                _cur_stream())
        _abi.check(rc, "unpack_rows")
        dev = tmp
    host = t.empty(tuple(dev.shape), dtype=t.float64, pin_memory=True)
    d2h.wait_stream(cur)
    with t.cuda.stream(d2h):
        host.copy_(dev, non_blocking=True)
        ev = t.cuda.Event()
        ev.record(d2h)
    dev.record_stream(d2h)
    return Pending(host, ev, layout.host_shape(n), complex_, cast)


# --------------------------------------------------------------------------
# optional per-kernel timing (bench.py): TIMERS[name] = [(start, end, bytes)]
# with CUDA events recorded on the launching stream around each launch
# --------------------------------------------------------------------------
TIMERS = None
TIMERS_ONLY = None      # set of launch names to time (None: all) -- bench.py


def _launch(name, alg_bytes, fn, *args):
    if TIMERS is None or (TIMERS_ONLY is not None and name not in TIMERS_ONLY):
        return fn(*args)
    t = torch()
    start, end = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    start.record()
    rc = fn(*args)
    end.record()
    TIMERS.setdefault(name, []).append((start, end, alg_bytes))
    return rc


# --------------------------------------------------------------------------
# plans (cached by their defining coefficients)
# --------------------------------------------------------------------------
class _Plan:
    _destroy = None

    def __init__(self):
        self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            if self.handle and self._destroy:
                getattr(_abi.load(), self._destroy)(self.handle)
        except Exception:
            pass


class _Cache:
    """LRU keyed by (current device, ...); producers may be iterated from several
    threads, so look-ups are serialised."""

    def __init__(self, size=32):
        import threading

        self.size, self.items, self.lock = size, OrderedDict(), threading.Lock()

    def get(self, key, make):
        dev = torch().cuda.current_device()
        key = (dev,) + key
        with self.lock:
            if key in self.items:
                self.items.move_to_end(key)
                return self.items[key]
            val = make()
            self.items[key] = val
            if len(self.items) > self.size:
                self.items.popitem(last=False)
            return val


_plans = _Cache()
_calib_cache = _Cache(64)


# Arithmetic of the FIR overlap-save transforms, the decimating polyphase filter
# and the Welch accumulation:
# "float64" (default: the reference computes and returns float64) or "float32"
# (opt-in: float64 samples in and out, FFTs in float32, ~1e-6 of the output peak;
# north_star's float32 tolerance is 1e-5).
COMPUTE = os.environ.get("OSZ_COMPUTE", "float64")


def set_compute(kind):
    """Select the arithmetic of the FFT-based FIR path, the decimating polyphase
    filter and the Welch accumulation: "float64" | "float32"."""
    global COMPUTE
    if kind not in ("float64", "float32"):
        raise ValueError("compute must be 'float64' or 'float32'")
    COMPUTE = kind


# Type of the SAMPLES in device memory and of the arrays handed back: "float64" (default:
# the reference returns float64 / complex128 for every input dtype, numerical.py:699) or
# "float32" (opt-in float32 I/O mode, SURVEY.md 8b / 8d: float32 in, float32 / complex64
# out, half the HBM and PCIe traffic; implies the float32 arithmetic of `set_compute`;
# IIR recurrences, their carried state and all sums stay float64).
IO = os.environ.get("OSZ_IO", "float64")
if IO == "float32":
    COMPUTE = "float32"


def set_io(kind):
    """"float64" | "float32": see IO.  Switching to float32 also selects the float32
    arithmetic; switching back restores float64 arithmetic."""
    global IO
    if kind not in ("float64", "float32"):
        raise ValueError("io must be 'float64' or 'float32'")
    IO = kind
    set_compute(kind)


def rows_dtype():
    t = torch()
    return t.float32 if IO == "float32" else t.float64


def to_f64(x):
    """float32 device rows -> float64 rows (operators without a float32 kernel)."""
    t = torch()
    if x.dtype == t.float64:
        return x
    rows, n = x.shape
    out = t.empty((rows, n), dtype=t.float64, device=x.device)
    ld_src = x.stride(0) if rows > 1 else max(n, 1)
    _abi.check(_abi.load().osz_widen_rows_f32_f64(_vp(x.data_ptr()), int(ld_src),
                                                  _vp(out.data_ptr()), n, rows, n, _cur_stream()),
               "widen")
    return out


def to_io(x, out=None):
    """float64 device rows -> rows of the I/O type (a no-op in float64 mode)."""
    t = torch()
    if x.dtype == rows_dtype() and out is None:
        return x
    rows, n = x.shape
    if out is None:
        out = t.empty((rows, n), dtype=rows_dtype(), device=x.device)
    if x.dtype == out.dtype:
        out.copy_(x)
        return out
    ld_src = x.stride(0) if rows > 1 else max(n, 1)
    ld_dst = out.stride(0) if rows > 1 else max(n, 1)
    _abi.check(_abi.load().osz_narrow_rows_f64_f32(_vp(x.data_ptr()), int(ld_src),
                                                   _vp(out.data_ptr()), int(ld_dst), rows, n,
                                                   _cur_stream()), "narrow")
    return out


class FirPlan(_Plan):
    _destroy = "osz_fir_plan_destroy"

    def __init__(self, taps, algo=_abi.FIR_AUTO):
        super().__init__()
        require_cuda()
        arr, ptr = _abi.as_double_array(taps)
        self.ntaps = int(arr.size)
        rc = _abi.load().osz_fir_plan_create(ctypes.byref(self.handle), ptr, self.ntaps, int(algo))
        _abi.check(rc, "fir_plan_create")
        self.algo = _abi.load().osz_fir_plan_algo(self.handle)

    @staticmethod
    def cached(taps, algo=_abi.FIR_AUTO):
        arr = np.ascontiguousarray(taps, dtype=np.float64)
        if algo == _abi.FIR_AUTO and COMPUTE == "float32" and 24 < arr.size <= 1025:
            algo = _abi.FIR_FFT_F32
        return _plans.get(("fir", arr.tobytes(), int(algo)), lambda: FirPlan(arr, algo))

    def run(self, xbuf, n_out, out=None):
        """xbuf: (rows, >= n_out + ntaps - 1) halo'd input.  Returns (rows, n_out)."""
        t = torch()
        rows = xbuf.shape[0]
        assert xbuf.shape[1] >= n_out + self.ntaps - 1
        if xbuf.dtype == t.float32:
            if out is None:
                out = empty_rows((rows, n_out))
            if self.algo != _abi.FIR_FFT_F32:
                # no float32 kernel for this tap count: float64 between a widen and a narrow
                y64 = self.run(to_f64(xbuf[:, :n_out + self.ntaps - 1]), n_out)
                return to_io(y64, out)
            xp, ldx = _rows_ptr(xbuf)
            yp, ldy = _rows_ptr(out)
            rc = _launch("fir", 8 * rows * n_out, _abi.load().osz_fir_exec_f32, self.handle, xp,
                         ldx, rows, n_out, yp, ldy, _cur_stream())
            _abi.check(rc, "fir_exec_f32")
            return out
        if out is None:
            out = empty((rows, n_out))
        xp, ldx = _rows_ptr(xbuf)
        yp, ldy = _rows_ptr(out)
        rc = _launch("fir", 16 * rows * n_out, _abi.load().osz_fir_exec_f64, self.handle, xp,
                     ldx, rows, n_out, yp, ldy, _cur_stream())
        _abi.check(rc, "fir_exec")
        return out


class SosPlan(_Plan):
    _destroy = "osz_sos_plan_destroy"

    def __init__(self, sos):
        super().__init__()
        require_cuda()
        arr, ptr = _abi.as_double_array(sos)
        assert arr.ndim == 2 and arr.shape[1] == 6
        self.nsec = int(arr.shape[0])
        rc = _abi.load().osz_sos_plan_create(ctypes.byref(self.handle), ptr, self.nsec)
        _abi.check(rc, "sos_plan_create")

    @staticmethod
    def cached(sos):
        arr = np.ascontiguousarray(sos, dtype=np.float64)
        return _plans.get(("sos", arr.tobytes()), lambda: SosPlan(arr))

    def run(self, x, state, reverse=False, want_output=True, out=None):
        """Filter x (rows, n) through the cascade; ``state`` (rows, nsec, 2) is
        updated in place.  Returns y, or None when only the state is wanted."""
        t = torch()
        rows, n = x.shape
        assert state.shape == (rows, self.nsec, 2) and state.is_contiguous()
        f32 = x.dtype == t.float32
        xp, ldx = _rows_ptr(x)
        if want_output:
            if out is None:
                out = t.empty((rows, n), dtype=x.dtype, device=x.device)
            assert out.dtype == x.dtype
            yp, ldy = _rows_ptr(out)
        else:
            out, yp, ldy = None, _vp(0), 0
        per = (4 if f32 else 8) * (2 if want_output else 1)
        fn = _abi.load().osz_sos_exec_f32 if f32 else _abi.load().osz_sos_exec_f64
        rc = _launch("sos" if want_output else "sos_state", per * rows * n, fn, self.handle, xp,
                     ldx, rows, n, int(bool(reverse)), _vp(state.data_ptr()), yp, ldy,
                     _cur_stream())
        _abi.check(rc, "sos_exec")
        return out

    def lookahead(self, zi, x, reverse=True):
        """State left by filtering x (rows, n) -- reversed: last sample first -- from
        ``zi * (first sample processed)``: the look-ahead pass of the forward-backward
        filters as one call.  Returns the (rows, nsec, 2) state."""
        rows, n = x.shape
        zarr, zptr = _abi.as_double_array(zi)
        assert zarr.shape == (self.nsec, 2)
        state = empty((rows, self.nsec, 2))
        f32 = x.dtype == torch().float32
        xp, ldx = _rows_ptr(x)
        fn = _abi.load().osz_sos_lookahead_f32 if f32 else _abi.load().osz_sos_lookahead_f64
        rc = _launch("sos_state", (4 if f32 else 8) * rows * n, fn, self.handle, zptr, xp, ldx,
                     rows, n, int(bool(reverse)), _vp(state.data_ptr()), _cur_stream())
        _abi.check(rc, "sos_lookahead")
        return state

    @property
    def has_weights(self):
        """True for cascades of one or two decaying sections: ``tail_state`` works."""
        if not hasattr(self, "_has_weights"):
            self._has_weights = bool(_abi.load().osz_sos_plan_has_weights(self.handle))
            self._settle = int(_abi.load().osz_sos_plan_settle(self.handle))
        return self._has_weights

    def tail_state(self, x, reverse=False):
        """State after filtering x (rows, n >= settle) from rest, as a weighted sum of
        the last ``settle`` samples processed (no recurrence, fully parallel)."""
        rows, n = x.shape
        state = empty((rows, self.nsec, 2))
        xp, ldx = _rows_ptr(x)
        rc = _launch("sos_state", 8 * rows * min(n, self._settle),
                     _abi.load().osz_sos_tail_state_f64, self.handle, xp, ldx, rows, n,
                     int(bool(reverse)), _vp(state.data_ptr()), _cur_stream())
        _abi.check(rc, "sos_tail_state")
        return state

    def state_from_sample(self, zi, x, sample):
        """state[r, s, :] = zi[s, :] * x[r, sample]."""
        t = torch()
        rows = x.shape[0]
        zarr, zptr = _abi.as_double_array(zi)
        assert zarr.shape == (self.nsec, 2)
        state = empty((rows, self.nsec, 2))
        xp, ldx = _rows_ptr(x)
        fn = (_abi.load().osz_sos_state_from_sample_f32 if x.dtype == t.float32
              else _abi.load().osz_sos_state_from_sample_f64)
        rc = fn(self.handle, zptr, xp, ldx, rows, int(sample), _vp(state.data_ptr()),
                _cur_stream())
        _abi.check(rc, "sos_state_from_sample")
        return state


class TfPlan(_Plan):
    """(b, a) filter of any order, transposed direct form II like scipy's lfilter."""
    _destroy = "osz_tf_plan_destroy"

    def __init__(self, b, a):
        super().__init__()
        require_cuda()
        barr, bptr = _abi.as_double_array(np.atleast_1d(b))
        aarr, aptr = _abi.as_double_array(np.atleast_1d(a))
        rc = _abi.load().osz_tf_plan_create(ctypes.byref(self.handle), bptr, len(barr), aptr,
                                            len(aarr))
        _abi.check(rc, "tf_plan_create")
        self.nstate = int(max(len(barr), len(aarr)) - 1)

    @staticmethod
    def cached(b, a):
        b = np.ascontiguousarray(np.atleast_1d(b), dtype=np.float64)
        a = np.ascontiguousarray(np.atleast_1d(a), dtype=np.float64)
        return _plans.get(("tf", b.tobytes(), a.tobytes()), lambda: TfPlan(b, a))

    def run(self, x, state, reverse=False, want_output=True, out=None):
        """Filter x (rows, n); ``state`` (rows, nstate) is updated in place."""
        rows, n = x.shape
        assert state.shape == (rows, self.nstate) and state.is_contiguous()
        if x.dtype == torch().float32:           # float32 I/O: float64 between widen and narrow
            y64 = self.run(to_f64(x), state, reverse=reverse, want_output=want_output)
            return to_io(y64, out) if want_output else None
        xp, ldx = _rows_ptr(x)
        if want_output:
            if out is None:
                out = empty((rows, n))
            yp, ldy = _rows_ptr(out)
        else:
            out, yp, ldy = None, _vp(0), 0
        rc = _launch("tf", 16 * rows * n, _abi.load().osz_tf_exec_f64, self.handle, xp, ldx, rows,
                     n, int(bool(reverse)), _vp(state.data_ptr()), yp, ldy, _cur_stream())
        _abi.check(rc, "tf_exec")
        return out

    def state_from_sample(self, zi, x, sample):
        """state[r, :] = zi * x[r, sample]."""
        rows = x.shape[0]
        zarr, zptr = _abi.as_double_array(zi)
        assert zarr.shape == (self.nstate,)
        state = empty((rows, self.nstate))
        if x.dtype == torch().float32:
            x = to_f64(x[:, int(sample):int(sample) + 1])
            sample = 0
        xp, ldx = _rows_ptr(x)
        rc = _abi.load().osz_tf_state_from_sample_f64(self.handle, zptr, xp, ldx, rows,
                                                      int(sample), _vp(state.data_ptr()),
                                                      _cur_stream())
        _abi.check(rc, "tf_state_from_sample")
        return state


class UpfirdnPlan(_Plan):
    _destroy = "osz_upfirdn_plan_destroy"

    KERNELS = ("auto", "polyphase", "mma", "general")

    def __init__(self, h, up, down, compute="float64", kernel="auto"):
        super().__init__()
        require_cuda()
        arr, ptr = _abi.as_double_array(h)
        self.ntaps, self.up, self.down = int(arr.size), int(up), int(down)
        rc = _abi.load().osz_upfirdn_plan_create(ctypes.byref(self.handle), ptr, self.ntaps,
                                                 self.up, self.down)
        _abi.check(rc, "upfirdn_plan_create")
        if kernel != "auto":
            _abi.check(_abi.load().osz_upfirdn_plan_set_kernel(self.handle,
                                                               self.KERNELS.index(kernel)),
                       "upfirdn_plan_set_kernel")
        # the float64 decimating kernel that will run: "mma" (FP64 tensor cores) |
        # "polyphase" (CUDA cores) | "general" (up > 1)
        self.kernel = self.KERNELS[_abi.load().osz_upfirdn_plan_kernel(self.handle)]
        if compute == "float32":
            _abi.check(_abi.load().osz_upfirdn_plan_set_compute(self.handle, 1),
                       "upfirdn_plan_set_compute")
        # float32 exists for the decimating kernel (up == 1) only
        self.compute = ("float64", "float32")[_abi.load().osz_upfirdn_plan_compute(self.handle)]

    @staticmethod
    def cached(h, up, down):
        arr = np.ascontiguousarray(h, dtype=np.float64)
        return _plans.get(("ufd", arr.tobytes(), int(up), int(down), COMPUTE),
                          lambda: UpfirdnPlan(arr, up, down, COMPUTE))

    def run(self, x, x_first, out_first, n_out, out=None):
        """x: (rows, m) holding global input samples x_first .. x_first+m-1.
        Returns global output samples out_first .. out_first+n_out-1."""
        rows, m = x.shape
        if x.dtype == torch().float32:
            if out is None:
                out = empty_rows((rows, n_out))
            if self.compute != "float32":
                # (up > 1, or a tile that does not fit): float64 between widen and narrow
                y64 = self.run(to_f64(x), x_first, out_first, n_out)
                return to_io(y64, out)
            xp, ldx = _rows_ptr(x)
            yp, ldy = _rows_ptr(out)
            rc = _launch("upfirdn", 4 * rows * (n_out * self.down // self.up + n_out),
                         _abi.load().osz_upfirdn_exec_f32, self.handle, xp, ldx, rows,
                         int(x_first), m, int(out_first), int(n_out), yp, ldy, _cur_stream())
            _abi.check(rc, "upfirdn_exec_f32")
            return out
        if out is None:
            out = empty((rows, n_out))
        xp, ldx = _rows_ptr(x)
        yp, ldy = _rows_ptr(out)
        rc = _launch("upfirdn", 8 * rows * (n_out * self.down // self.up + n_out),
                     _abi.load().osz_upfirdn_exec_f64, self.handle, xp, ldx, rows, int(x_first),
                     m, int(out_first), int(n_out), yp, ldy, _cur_stream())
        _abi.check(rc, "upfirdn_exec")
        return out


# --------------------------------------------------------------------------
# fused IIR pass + decimating FIR (csrc/sosdec.cu)
# --------------------------------------------------------------------------
def sosdec_spans(sos_plan, ufd_plan, rows, n):
    """Time spans per row the fused kernel would use; 0 = cannot run fused."""
    return int(_abi.load().osz_sosdec_spans(sos_plan.handle, ufd_plan.handle, int(rows), int(n)))


def sosdec_exec(sos_plan, ufd_plan, x, reverse, state, nspan, first, out, out_first):
    """Run the pass over x (rows, n) and decimate its output on chip.  Writes the
    outputs whose window lies inside one span into ``out`` (column c = global output
    out_first + c) and returns the (rows, nspan, 2, K-1) edge samples."""
    rows, n = x.shape
    k1 = ufd_plan.ntaps - 1
    edges = empty((rows, int(nspan), 2, k1))
    xp, ldx = _rows_ptr(x)
    yp, ldy = _rows_ptr(out)
    rc = _launch("sos_dec", 8 * rows * n + 8 * rows * out.shape[1],
                 _abi.load().osz_sosdec_exec_f64, sos_plan.handle, ufd_plan.handle, xp, ldx, rows,
                 n, int(bool(reverse)), _vp(state.data_ptr()), int(nspan), int(first), yp, ldy,
                 int(out_first), int(out.shape[1]), _vp(edges.data_ptr()), _cur_stream())
    _abi.check(rc, "sosdec_exec")
    return edges


def sosdec_boundary(ufd_plan, edges, reverse, prev_tail, has_end, n, first, out, out_first, j_min,
                    j_max):
    """The outputs straddling the chunk's start, span boundaries and (has_end) the
    recording's end.  prev_tail: (rows, K-1) view (row stride free) or None."""
    rows, nspan = edges.shape[0], edges.shape[1]
    yp, ldy = _rows_ptr(out)
    if prev_tail is not None:
        assert prev_tail.stride(1) == 1
        pt, pld = _vp(prev_tail.data_ptr()), int(prev_tail.stride(0))
    else:
        pt, pld = _vp(0), 0
    rc = _launch("sos_dec_edges", 0, _abi.load().osz_sosdec_boundary_f64, ufd_plan.handle,
                 _vp(edges.data_ptr()), int(nspan), int(bool(reverse)), pt, pld, int(bool(has_end)),
                 rows, int(n), int(first), yp, ldy, int(out_first), int(out.shape[1]), int(j_min),
                 int(j_max), _cur_stream())
    _abi.check(rc, "sosdec_boundary")


class SpecPlan(_Plan):
    _destroy = "osz_spec_plan_destroy"

    def __init__(self, nfft, stride, window, detrend, norm, compute="float64"):
        super().__init__()
        require_cuda()
        arr, ptr = _abi.as_double_array(window)
        assert arr.size == nfft
        self.nfft, self.stride, self.nfreq = int(nfft), int(stride), int(nfft) // 2 + 1
        rc = _abi.load().osz_spec_plan_create(ctypes.byref(self.handle), self.nfft, self.stride,
                                              ptr, _abi.DETREND[detrend], float(norm))
        _abi.check(rc, "spec_plan_create")
        self.path = _abi.load().osz_spec_plan_path(self.handle)
        if compute == "float32":
            rc = _abi.load().osz_spec_plan_set_compute(self.handle, 1, ptr)
            _abi.check(rc, "spec_plan_set_compute")
        # what the plan's Welch accumulation really computes in (float32 exists for
        # power-of-two nfft 512 .. 4096 only)
        self.compute = ("float64", "float32")[_abi.load().osz_spec_plan_compute(self.handle)]

    @staticmethod
    def cached(nfft, stride, window, detrend, norm):
        arr = np.ascontiguousarray(window, dtype=np.float64)
        key = ("spec", int(nfft), int(stride), arr.tobytes(), str(detrend), float(norm), COMPUTE)
        return _plans.get(key, lambda: SpecPlan(nfft, stride, arr, detrend, norm, COMPUTE))

    def nseg_available(self, width):
        return (width - self.nfft) // self.stride + 1 if width >= self.nfft else 0

    def welch_accum(self, x, nseg, psd_sum):
        rows = x.shape[0]
        assert x.shape[1] >= (nseg - 1) * self.stride + self.nfft
        assert psd_sum.shape == (rows, self.nfreq) and psd_sum.is_contiguous()
        f32 = x.dtype == torch().float32
        if f32 and self.compute != "float32":
            x, f32 = to_f64(x), False            # no float32 kernel for this nfft
        xp, ldx = _rows_ptr(x)
        fn = _abi.load().osz_welch_accum_f32 if f32 else _abi.load().osz_welch_accum_f64
        rc = _launch("welch", (4 if f32 else 8) * rows * int(nseg) * self.stride, fn, self.handle,
                     xp, ldx, rows, int(nseg), _vp(psd_sum.data_ptr()), self.nfreq, _cur_stream())
        _abi.check(rc, "welch_accum")

    def segments(self, x, nseg, complex_):
        """Per-segment modified DFT (complex_) or periodogram: (nseg, rows, nfreq[, 2])."""
        t = torch()
        rows = x.shape[0]
        assert x.shape[1] >= (nseg - 1) * self.stride + self.nfft
        shape = (nseg, rows, self.nfreq) + ((2,) if complex_ else ())
        out = empty(shape)
        x = to_f64(x)          # (float32 I/O: the per-segment kernels read float64 samples)
        xp, ldx = _rows_ptr(x)
        fn = _abi.load().osz_stft_f64 if complex_ else _abi.load().osz_periodogram_f64
        rc = fn(self.handle, xp, ldx, rows, int(nseg), _vp(out.data_ptr()), _cur_stream())
        _abi.check(rc, "stft" if complex_ else "periodogram")
        return out


def spec_prepare(x, n, nfft, window, detrend):
    """(rows, nfft) device rows: the first n samples of x detrended and
    windowed, then zeros (periodogram / modified_dft with nfft > n)."""
    rows = x.shape[0]
    x = to_f64(x)
    w = from_host(np.ascontiguousarray(window, dtype=np.float64))
    out = empty((rows, int(nfft)))
    xp, ldx = _rows_ptr(x)
    op, ldo = _rows_ptr(out)
    rc = _abi.load().osz_spec_prepare_f64(xp, ldx, rows, int(n), int(nfft), _vp(w.data_ptr()),
                                          _abi.DETREND[detrend], op, ldo, _cur_stream())
    _abi.check(rc, "spec_prepare")
    return out


# --------------------------------------------------------------------------
# producer tools (csrc/protools.cu)
# --------------------------------------------------------------------------
def take_cols(x, idx):
    """x[:, idx] for device rows ``(rows, n)`` and ascending host int64 positions
    ``idx`` (np.flatnonzero of a mask chunk): the device compaction of
    MaskedProducer (reference core/producer.py:436-437)."""
    t = require_cuda()
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    rows = x.shape[0]
    if x.dtype == t.float32:                 # float32 I/O: float64 between widen and narrow
        return to_io(take_cols(to_f64(x), idx))
    out = empty((rows, int(idx.size)))
    if idx.size == 0 or rows == 0:
        return out
    if int(idx[0]) < 0 or int(idx[-1]) >= x.shape[1]:
        # np.take's error for a mask that is True beyond the end of the data
        raise IndexError("index {} is out of bounds for axis with size {}".format(
            int(idx[-1]), x.shape[1]))
    # pinned staging + asynchronous copy on the compute stream: a pageable `.to()` would
    # block the host until the copy is done and serialise the chunk pipeline
    pin = t.empty(int(idx.size), dtype=t.int64, pin_memory=True)
    pin.numpy()[:] = idx
    idx_dev = pin.to(DEVICE, non_blocking=True)
    xp, ldx = _rows_ptr(x)
    yp, ldy = _rows_ptr(out)
    rc = _launch("take_cols", 16 * rows * int(idx.size), _abi.load().osz_take_cols_f64, xp, ldx,
                 rows, _vp(idx_dev.data_ptr()), int(idx.size), yp, ldy, _cur_stream())
    _abi.check(rc, "take_cols")
    return out


class RowMoments:
    """Running per-row sums for protools.mean / protools.std along the
    production axis: ``add`` folds one chunk in on the device, ``result`` brings
    back (sum of n * chunk mean, sum of n * chunk mean of squares, sum of n)."""

    def __init__(self, rows, ignore_nan=True):
        require_cuda()
        self.rows, self.ignore_nan = int(rows), bool(ignore_nan)
        self.acc = zeros((self.rows, 3))
        self.scratch = empty((self.rows, _abi.load().osz_row_moments_slots() * 3))

    def add(self, x):
        assert x.shape[0] == self.rows
        x = to_f64(x)
        xp, ldx = _rows_ptr(x)
        rc = _launch("row_moments", 8 * self.rows * x.shape[1], _abi.load().osz_row_moments_f64,
                     xp, ldx, self.rows, int(x.shape[1]), int(self.ignore_nan),
                     _vp(self.acc.data_ptr()), _vp(self.scratch.data_ptr()), _cur_stream())
        _abi.check(rc, "row_moments")

    def result(self):
        return self.acc.cpu().numpy()


def row_standardize(x, mean_dev, std_dev, out=None):
    """(x - mean[r]) / std[r] per device row (protools.standardize, production axis)."""
    rows, n = x.shape
    if x.dtype == torch().float32:
        return to_io(row_standardize(to_f64(x), mean_dev, std_dev), out)
    if out is None:
        out = empty((rows, n))
    xp, ldx = _rows_ptr(x)
    yp, ldy = _rows_ptr(out)
    rc = _launch("standardize", 16 * rows * n, _abi.load().osz_row_standardize_f64, xp, ldx, rows,
                 int(n), _vp(mean_dev.data_ptr()), _vp(std_dev.data_ptr()), yp, ldy, _cur_stream())
    _abi.check(rc, "row_standardize")
    return out


def col_moments(x, ignore_nan=True, want="mean"):
    """Per-sample statistics ACROSS the rows of a device chunk: ``want`` =
    "mean" | "std" -> (1, n) rows, "standardize" -> (rows, n)."""
    rows, n = x.shape
    if x.dtype == torch().float32:
        return to_io(col_moments(to_f64(x), ignore_nan, want))
    xp, ldx = _rows_ptr(x)
    null = _vp(0)
    if want == "standardize":
        out = empty((rows, n))
        yp, ldy = _rows_ptr(out)
        args = (null, null, yp, ldy)
    else:
        out = empty((1, n))
        args = ((_vp(out.data_ptr()), null) if want == "mean" else (null, _vp(out.data_ptr()))) \
            + (null, 0)
    rc = _launch("col_moments", 8 * rows * n, _abi.load().osz_col_moments_f64, xp, ldx, rows,
                 int(n), int(bool(ignore_nan)), *args, _cur_stream())
    _abi.check(rc, "col_moments")
    return out


def zip_complex(re, im):
    """(rows, n) real and imaginary device rows -> (rows, n, 2) complex128 rows."""
    rows, n = re.shape
    assert im.shape == re.shape
    re, im = to_f64(re), to_f64(im)
    out = empty((rows, n, 2))
    rp, ldr = _rows_ptr(re)
    ip, ldi = _rows_ptr(im)
    rc = _launch("zip_complex", 32 * rows * n, _abi.load().osz_zip_complex_f64, rp, ldr, ip, ldi,
                 rows, int(n), _vp(out.data_ptr()), _cur_stream())
    _abi.check(rc, "zip_complex")
    return out


def ceil_div(a, b):
    return -(-int(a) // int(b))


def cat_time(parts):
    """Concatenate device row blocks along time."""
    parts = [p for p in parts if p is not None and p.shape[1] > 0]
    if len(parts) == 1:
        return parts[0]
    return torch().cat(parts, dim=1)


def zeros(shape):
    """Zero-filled float64 device tensor (states, sums)."""
    t = require_cuda()
    return t.zeros(tuple(shape), dtype=t.float64, device=DEVICE)


def empty(shape):
    """float64 device tensor (states, sums, spectra)."""
    t = require_cuda()
    return t.empty(tuple(shape), dtype=t.float64, device=DEVICE)


def empty_rows(shape):
    """Sample rows: float64, or float32 in the float32 I/O mode."""
    t = require_cuda()
    return t.empty(tuple(shape), dtype=rows_dtype(), device=DEVICE)


def zeros_rows(rows, n):
    t = require_cuda()
    return t.zeros((rows, n), dtype=rows_dtype(), device=DEVICE)


def record_event():
    """Event on the current (compute) stream, or None off-GPU (test stand-ins)."""
    t = torch()
    if DEVICE != "cuda":
        return None
    ev = t.cuda.Event()
    ev.record()
    return ev


def from_host(arr):
    """Small host ndarray -> device tensor (synchronous; coefficients / states)."""
    t = require_cuda()
    return t.from_numpy(np.array(arr, dtype=np.float64, order="C", copy=True)).to(DEVICE)


def isfinite_number(x):
    return isinstance(x, (int, float)) and math.isfinite(x)
