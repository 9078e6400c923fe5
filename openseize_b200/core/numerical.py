"""GPU generating functions with the names, positional order and chunk
semantics of the reference's ``openseize.core.numerical`` (the hot path,
reference core/numerical.py:158-1087).  The L4 operators build
``functools.partial(nm.<genfunc>, pro, ...)`` and wrap it in a producer exactly
as the reference does; nothing touches the GPU until the producer is iterated.

Every generating function ``f`` has a device twin ``f.device`` that yields
device row blocks instead of host ndarrays.  When the input producer of a GPU
operator is itself a producer over one of these functions, the chain is run on
the device end to end (``device_chunks``): one pull of each upstream chunk, no
host round trips between stages -- the reference re-executes upstream stages
once per downstream iterator; results are identical.

There is no CPU arithmetic in this module: without a CUDA device every
generator raises on first iteration.
"""

import functools

import numpy as np
import scipy.signal as sps

from openseize_b200.core import device as dv
from openseize_b200.core.arraytools import normalize_axis
from openseize_b200.core.producer import (DeviceProducer, GenProducer, MaskedProducer, Producer,
                                           producer)

# ---------------------------------------------------------------------------
# shape helpers (pure index arithmetic, reference core/numerical.py:19-155)
# ---------------------------------------------------------------------------


def optimal_nffts(arr):
    """The reference's block-size heuristic (numerical.py:19-38).  Kept for
    API parity; the GPU kernels pick their own block size (results do not
    depend on it)."""
    return int(8 * 2 ** np.ceil(np.log2(len(arr))))


def convolved_shape(shape1, shape2, mode, axis):
    """Shape of the convolution of two arrays along ``axis`` (numerical.py:41-73)."""
    m, n = shape1[axis], shape2[axis if len(shape2) > 1 else 0]
    p, q = max(m, n), min(m, n)
    out = list(shape1 if len(shape1) >= len(shape2) else shape2)
    out[axis] = {"full": m + n - 1, "same": p, "valid": m + n - 1 - 2 * (q - 1)}[mode]
    return tuple(out)


def _mode_cuts(ntaps, mode):
    """Samples of the full convolution dropped on the left / right by the numpy
    convolve modes (numerical.py:143-150)."""
    if mode == "full":
        return 0, 0
    if mode == "same":
        return (ntaps - 1) // 2, int(np.ceil((ntaps - 1) / 2))
    if mode == "valid":
        return ntaps - 1, ntaps - 1
    raise ValueError("mode must be one of 'full', 'same', 'valid', got {!r}".format(mode))


# ---------------------------------------------------------------------------
# device chunk streams
# ---------------------------------------------------------------------------
class _DeviceFifo:
    """Device twin of FIFOArray: collects row blocks, releases fixed-size ones."""

    def __init__(self):
        self.parts, self.size = [], 0

    def put(self, block):
        if block.shape[1]:
            self.parts.append(block)
            self.size += block.shape[1]

    def get(self, n):
        n = min(n, self.size)
        if self.parts and self.parts[0].shape[1] == n:
            self.size -= n
            return self.parts.pop(0)
        buf = dv.cat_time(self.parts)
        out, rest = buf[:, :n], buf[:, n:]
        self.parts = [rest] if rest.shape[1] else []
        self.size -= n
        return out


def _regrid_device(blocks, cs):
    """Re-block a stream of device rows to ``cs`` samples (last one shorter)."""
    fifo = _DeviceFifo()
    for block in blocks:
        fifo.put(block)
        while fifo.size >= cs:
            yield fifo.get(cs)
    if fifo.size:
        yield fifo.get(cs)


def _device_twin(pro):
    """(device generating function, args, kwargs) if ``pro`` is a producer over
    one of this module's GPU generating functions, else None."""
    if not isinstance(pro, GenProducer) or pro.kwargs:
        return None
    func, args, kwargs = pro.data, (), {}
    if isinstance(func, functools.partial):
        func, args, kwargs = func.func, func.args, func.keywords
    twin = getattr(func, "device", None)
    return (twin, args, kwargs) if twin is not None else None


_HOST_BLOCK_BYTES = 32 << 20


def device_chunks(pro, axis, regrid=True, alloc=None):
    """Yield ``pro``'s chunks as device rows ``(rows, chunk)``.

    Host producers are streamed through pinned memory with the copy of chunk
    k+1 issued while chunk k's kernels run; producers over GPU generating
    functions are consumed on the device.

    regrid: re-block a device upstream to the chunk grid of ``pro.chunksize``
        (needed by the chunk-dependent forward-backward filters); stages whose
        result does not depend on the grid take the upstream blocks as they come.
    alloc: ``alloc(rows, n) -> (rows, n) device view`` -- where the consumer
        wants the next block written (its halo'd staging ring), so neither an
        upload nor an upstream kernel needs a second copy.
    """
    dv.require_cuda()
    layout = dv.Layout(pro.shape, axis)
    if isinstance(pro, DeviceProducer):
        source = pro.device_iter()
    elif isinstance(pro, MaskedProducer) and pro.axis == layout.axis:
        source = _masked_device(pro)
    else:
        twin = _device_twin(pro)
        if twin is None:
            # A consumer that does not depend on the chunk grid (regrid=False) takes
            # arrays and readers in blocks of >= ~32 MB: psd pulls fs-sized chunks
            # (spectra/estimators.py:141), and a thousand 2 MB uploads cost more in
            # launches than in PCIe time.  Results do not change.
            step = int(pro.chunksize)
            if not regrid and hasattr(pro, "blocks"):
                per_chunk = max(layout.rows * step * 8, 1)
                step *= max(1, -(-_HOST_BLOCK_BYTES // per_chunk))
            # EDF readers hand over their int16 records: a quarter of the PCIe
            # traffic, calibrated on the device (file_io/edf.py)
            raw = pro.iter_raw(step) if hasattr(pro, "iter_raw") else None
            if raw is None:
                raw = pro.blocks(step) if hasattr(pro, "blocks") else pro
            for arr in raw:
                yield dv.upload(arr, layout, alloc)
            return
        func, args, kwargs = twin
        # _free: the consumer does not care where the upstream cuts its blocks
        source = func(*args, **dict(kwargs, _out=alloc if not regrid else None,
                                    _free=not regrid))
    if not regrid:
        yield from source
        return
    fifo, cs = _DeviceFifo(), int(pro.chunksize)
    for block in source:
        fifo.put(block)
        while fifo.size >= cs:
            yield fifo.get(cs)
    if fifo.size:
        yield fifo.get(cs)


def _masked_device(pro):
    """MaskedProducer on the device (reference core/producer.py:427-444): the
    inner producer's chunks -- uploaded, or handed over in HBM when it is a chain
    of GPU operators -- are compacted by ``take_cols`` with the positions of the
    mask chunk that pairs with them; chunks without a kept sample are skipped and
    production stops when the producer or the mask runs out (``zip``).  The
    caller re-blocks the stream to the chunk size as the reference's FIFO does."""
    inner = pro.data
    for block, keep in zip(device_chunks(inner, pro.axis, regrid=True), pro.mask):
        idx = np.flatnonzero(keep)
        if idx.size:
            yield dv.take_cols(block, idx)


def _new_rows(out, rows, n):
    """Output block of a stage: in the consumer's staging ring when it asked."""
    return out(rows, n) if out is not None else dv.empty_rows((rows, n))


class _TimeRing:
    """Device staging for stages that need bounded history along time (FIR
    halo, resampler reach, Welch overlap): blocks are appended at the write
    position -- ideally written there directly by the upstream kernel or the
    H2D copy through ``alloc`` -- and the stage reads ``window()``, the live
    span, as one contiguous halo'd view.  Replaces the concat-on-put /
    split-on-get of the reference's FIFOArray (core/queues.py:46-70)."""

    def __init__(self, rows, capacity=0):
        self.rows, self.buf, self.start, self.pos, self._pending = rows, None, 0, 0, None
        self.ready_event = None      # compute-stream point after which the buffer is ours
        # samples the consumer lets pile up before it reads (a stage that batches small
        # chunks): the buffer is sized for that once instead of being re-based -- with a
        # copy of everything live -- every other block
        self.capacity = int(capacity)

    def _reserve(self, n):
        live = self.pos - self.start
        if self.buf is None or self.pos + n > self.buf.shape[1]:
            # the write position and the row pitch stay multiples of 16 samples (as long
            # as the blocks are): upstream kernels that move tiles with the TMA need
            # 16-byte aligned rows (csrc/sos_tile.cuh)
            lead = -live % 16
            width = lead + max(live + 2 * n, self.capacity + 2 * n) + 64
            new = dv.empty_rows((self.rows, width + (-width % 16)))
            if live:
                new[:, lead:lead + live].copy_(self.buf[:, self.start:self.pos])
            self.buf, self.start, self.pos = new, lead, lead + live
            self.ready_event = dv.record_event()

    def __call__(self, rows, n):
        return self.alloc(rows, n)

    def alloc(self, rows, n):
        assert rows == self.rows
        self._reserve(n)
        self._pending = self.buf[:, self.pos:self.pos + n]
        return self._pending

    def _in_buffer(self, block):
        if self.buf is None:
            return False
        lo = self.buf.data_ptr()
        return lo <= block.data_ptr() < lo + self.buf.numel() * self.buf.element_size()

    def push(self, block):
        n = block.shape[1]
        pend, self._pending = self._pending, None
        in_place = (pend is not None and block.data_ptr() == pend.data_ptr()
                    and tuple(block.shape) == tuple(pend.shape)
                    and block.stride() == pend.stride())
        if not in_place:
            if self._in_buffer(block):
                block = block.clone()        # a cut of the pending region: copy would overlap
            self._reserve(n)
            self.buf[:, self.pos:self.pos + n].copy_(block)
        self.pos += n

    def push_zeros(self, n):
        if n > 0:
            self._reserve(n)
            self.buf[:, self.pos:self.pos + n].zero_()
            self.pos += n

    @property
    def size(self):
        return self.pos - self.start

    def window(self):
        return self.buf[:, self.start:self.pos]

    def drop(self, k):
        self.start += int(k)


def _to_host(device_gen, layout, complex_=False):
    """Drain a device generator to host ndarrays, one chunk behind the kernels
    so the D2H copy of chunk k overlaps the compute of chunk k+1."""
    pending = None
    for block in device_gen:
        nxt = dv.download(block, layout, complex_)
        if pending is not None:
            yield pending.get()
        pending = nxt
    if pending is not None:
        yield pending.get()


def _gpu_genfunc(device_func, out_layout, regrid=False):
    """Build the host generating function of a device generating function.
    ``out_layout(*args, **kwargs)`` gives the Layout of the yielded arrays.
    ``regrid``: cut the yielded blocks to the input producer's chunk grid (only
    for functions whose raw yield lengths are not part of the reference's contract)."""

    def decorate(host_stub):
        @functools.wraps(host_stub)
        def genfunc(*args, **kwargs):
            layout = out_layout(*args, **kwargs)
            blocks = device_func(*args, **kwargs)
            # The host consumer (GenProducer) re-chunks what we yield to the chunk size
            # of the operator call, which is the input producer's chunk size.  Cutting
            # the blocks to that grid on the DEVICE (HBM copies) lets them pass through
            # the host FIFO without the per-chunk np.concatenate that a stream of
            # off-grid blocks costs (FIR 'same' shortens its first block by the left cut,
            # so every later block would straddle: 4 ms per 32 MB chunk).
            cs = int(getattr(args[0], "chunksize", 0) or 0) if args else 0
            if regrid and cs > 0:
                blocks = _regrid_device(blocks, cs)
            yield from _to_host(blocks, layout)

        genfunc.device = device_func
        return genfunc

    return decorate


def _layout_of(pro, axis):
    return dv.Layout(pro.shape, axis)


# ---------------------------------------------------------------------------
# FIR  (reference core/numerical.py:158-298)
# ---------------------------------------------------------------------------
def _oaconvolve_device(pro, window, axis, mode, nfft_factor=32, _out=None, _free=False):
    dv.require_cuda()
    window = np.asarray(window, dtype=np.float64)
    ntaps, nsamp = len(window), pro.shape[axis]
    if nsamp < ntaps:
        raise ValueError("oaconvolve: data length {} along axis is shorter than the {} taps"
                         .format(nsamp, ntaps))
    left, right = _mode_cuts(ntaps, mode)
    last_kept = nsamp + ntaps - 1 - right          # exclusive, in full-convolution index
    plan = dv.FirPlan.cached(window)
    rows = _layout_of(pro, axis).rows
    ring = _TimeRing(rows)
    ring.push_zeros(ntaps - 1)                      # the reference's zero overlap (:221-223)
    pos = 0                                         # full-convolution index of the next output
    seen, flushed = 0, False
    for chunk in device_chunks(pro, axis, regrid=False, alloc=ring):
        n = chunk.shape[1]
        ring.push(chunk)
        seen += n
        final = seen >= nsamp
        if final:
            ring.push_zeros(ntaps - 1)              # flush the tail (:285-298)
            flushed = True
        n_out = n + (ntaps - 1 if final else 0)
        y = plan.run(ring.window(), n_out, out=_new_rows(_out, rows, n_out))
        ring.drop(n_out)                            # keep the last ntaps-1 samples as halo
        lo = max(left - pos, 0)
        hi = min(pos + n_out, last_kept) - pos
        pos += n_out
        if hi > lo:
            yield y if (lo == 0 and hi == n_out) else y[:, lo:hi]
        if final:
            break
    if not flushed and seen > 0:
        # The source ran out before its declared length (a generator with a quirky
        # shape, a masked stream cut short): the reference flushes when the iterator
        # ends (:285-298), so the tail belongs to the samples that did arrive.
        last_kept = seen + ntaps - 1 - right
        ring.push_zeros(ntaps - 1)
        n_out = ntaps - 1
        y = plan.run(ring.window(), n_out, out=_new_rows(_out, rows, n_out))
        lo = max(left - pos, 0)
        hi = min(pos + n_out, last_kept) - pos
        if hi > lo:
            yield y[:, lo:hi]


def _oaconvolve_layout(pro, window, axis, mode, nfft_factor=32):
    return _layout_of(pro, axis)


@_gpu_genfunc(_oaconvolve_device, _oaconvolve_layout, regrid=True)
def oaconvolve(pro, window, axis, mode, nfft_factor=32):
    """Convolve every 1-D slice of a producer along ``axis`` with ``window``;
    numpy convolve modes.  ``nfft_factor`` is accepted for signature parity and
    ignored (the GPU block size does not change the result).

    Yields one block per input chunk (the reference yields FFT-step sized
    blocks; the concatenation is identical)."""


# ---------------------------------------------------------------------------
# IIR  (reference core/numerical.py:301-520)
# ---------------------------------------------------------------------------
_MAX_SEC = 16
_LOOK_WEIGHTS = __import__("os").environ.get("OSZ_LOOK_WEIGHTS", "0") == "1"


def _sos_groups(sos):
    sos = np.atleast_2d(np.asarray(sos, dtype=np.float64))
    return [dv.SosPlan.cached(sos[i:i + _MAX_SEC]) for i in range(0, sos.shape[0], _MAX_SEC)]


def _cascade_transition(sos):
    """One-step zero-input transition of a DF2T biquad cascade (state order:
    z0, z1 of section 0, then section 1, ...): column j is the state after one
    zero input sample from unit state j; section s+1's input is section s's
    output (recurrence of reference numerical.py:334 / scipy's sosfilt)."""
    sos = np.atleast_2d(np.asarray(sos, dtype=np.longdouble))
    nsec = sos.shape[0]
    T = np.zeros((2 * nsec, 2 * nsec), dtype=np.longdouble)
    for j in range(2 * nsec):
        xin = np.longdouble(0)
        for s in range(nsec):
            b0, b1, b2, a0, a1, a2 = sos[s]
            b0, b1, b2, a1, a2 = b0 / a0, b1 / a0, b2 / a0, a1 / a0, a2 / a0
            z0 = np.longdouble(j == 2 * s)
            z1 = np.longdouble(j == 2 * s + 1)
            y = b0 * xin + z0
            T[2 * s, j] = b1 * xin - a1 * y + z1
            T[2 * s + 1, j] = b2 * xin - a2 * y
            xin = y
    return T


def _settle_from_transition(T, tol=1e-18):
    """Smallest n (plus a margin) with ||T^n||_inf < tol, found on the ladder
    T^(2^k) in long double and refined bit by bit -- exact in the pole
    multiplicities (m coinciding poles decay like n^(m-1) r^n, which a bound
    from the largest pole radius alone misses).  None when the norm never gets
    there (a pole on or outside the unit circle)."""
    T = np.asarray(T, dtype=np.longdouble)
    if T.size == 0:
        return 0

    def norm(A):
        return float(np.max(np.sum(np.abs(A), axis=1)))

    ladder = [T]
    with np.errstate(over="ignore", invalid="ignore"):
        while True:
            v = norm(ladder[-1])
            if not np.isfinite(v) or v > 1e300:
                return None
            if v < tol:
                break
            if len(ladder) > 40:
                return None
            ladder.append(ladder[-1] @ ladder[-1])
        k = len(ladder) - 1
        if k == 0:
            return 2 * T.shape[0]
        acc, steps = ladder[k - 1], 1 << (k - 1)
        for j in range(k - 2, -1, -1):
            cand = acc @ ladder[j]
            if norm(cand) >= tol:
                acc, steps = cand, steps + (1 << j)
    return steps + steps // 16 + 64


class _Cascade:
    """A biquad cascade of any length as groups of <= 16 sections."""

    def __init__(self, sos):
        self.sos = np.atleast_2d(np.asarray(sos, dtype=np.float64))
        self.plans = _sos_groups(self.sos)
        self.nsec = self.sos.shape[0]
        self.settle = self._settle_samples()

    def _settle_samples(self):
        """Samples after which the cascade has forgotten its initial state (to
        1e-18 of it); None if a pole sits on or outside the unit circle."""
        return _settle_from_transition(_cascade_transition(self.sos))

    def split_state(self, state):
        """(rows, nsec, 2) -> contiguous per-group states."""
        out, s = [], 0
        for p in self.plans:
            out.append(state[:, s:s + p.nsec, :].contiguous())
            s += p.nsec
        return out

    def zero_state(self, rows):
        return [dv.zeros((rows, p.nsec, 2)) for p in self.plans]

    def state_from_sample(self, zi, x, sample):
        """Reference: zi * x[..., sample] per section (numerical.py:385,399,410).
        Section s+1 sees the steady-state OUTPUT of section s, which
        scipy.signal.sosfilt_zi has already folded into zi."""
        out, s = [], 0
        for p in self.plans:
            out.append(p.state_from_sample(zi[s:s + p.nsec], x, sample))
            s += p.nsec
        return out

    def look_ahead(self, zi, fwd):
        """State the backward pass of the PREVIOUS chunk starts from: the state left by
        filtering this forward chunk backwards from zi * (its last sample) (reference
        numerical.py:397-399, :508-509).  A stable cascade has forgotten where it
        started after `settle` samples, so only the first `settle` samples of the chunk
        matter (to ~1e-18): they are either re-filtered by a state-only pass, or -- for
        one or two sections -- folded into the state by one weighted sum."""
        m = fwd.shape[1]
        if self.settle is not None and self.settle < m:
            m = self.settle
            plan = self.plans[0]
            # (opt-in: one launch instead of two, but it reads two weights per sample --
            #  0.12 ms against 0.05 ms for the state-only scan at 256 rows on B200)
            if (_LOOK_WEIGHTS and len(self.plans) == 1
                    and getattr(plan, "has_weights", False)):
                return [plan.tail_state(fwd[:, :m], reverse=True)]
        if len(self.plans) == 1:
            # one call: start state zi * fwd[m - 1] and the reversed state-only pass
            return [self.plans[0].lookahead(zi[:self.plans[0].nsec], fwd[:, :m])]
        look = self.state_from_sample(zi, fwd, m - 1)
        self.run(fwd[:, :m], look, reverse=True, want_output=False)
        return look

    def run(self, x, states, reverse=False, want_output=True, out=None):
        y, last = x, len(self.plans) - 1
        for i, (p, st) in enumerate(zip(self.plans, states)):
            need = want_output or i < last
            dst = _new_rows(out, x.shape[0], x.shape[1]) if (need and i == last) else None
            y = p.run(y, st, reverse=reverse, want_output=need, out=dst)
        return y if want_output else None


def _zi_to_rows(zi, layout, nsec):
    """Reference zi layout (nsec, ..., 2, ...) -> (rows, nsec, 2) device tensor."""
    zi = np.asarray(zi, dtype=np.float64)
    full = (nsec,) + layout.host_shape(2)
    zi = np.broadcast_to(zi, full).reshape(nsec, layout.outer, 2, layout.inner)
    rows = np.ascontiguousarray(np.transpose(zi, (1, 3, 0, 2))).reshape(layout.rows, nsec, 2)
    return dv.from_host(rows)


def _sosfilt_device(pro, sos, axis, zi=None, _out=None, _free=False):
    dv.require_cuda()
    layout = _layout_of(pro, axis)
    cascade = _Cascade(sos)
    if zi is None:
        states = cascade.zero_state(layout.rows)
    else:
        states = cascade.split_state(_zi_to_rows(zi, layout, cascade.nsec))
    for chunk in device_chunks(pro, axis, regrid=False):
        yield cascade.run(chunk, states, out=_out)


def _same_layout(pro, *args, **kwargs):
    axis = args[1] if len(args) > 1 else kwargs["axis"]
    return _layout_of(pro, axis)


@_gpu_genfunc(_sosfilt_device, _same_layout)
def sosfilt(pro, sos, axis, zi=None):
    """Forward second-order-section filter with the delay registers carried
    from chunk to chunk (reference numerical.py:301-335)."""


def _filtfilt_device(pro, cascade, zi, axis, _out=None, _fwd_states=None, _drop_last=False):
    """Shared forward-backward driver (reference numerical.py:338-411,449-520):
    one global forward pass; each chunk's backward pass starts from the state
    left by filtering the NEXT forward chunk backwards from zi * its last
    sample; the last chunk starts from zi * its own last sample.

    Time shards (sharding.iir_time_sharded) run this on a span of the recording:
    ``_fwd_states`` is the forward state entering the span (instead of
    zi * first sample) and ``_drop_last`` marks the span's last chunk as the
    look-ahead chunk borrowed from the next span -- it is filtered forward but
    its own backward pass belongs to the next rank."""
    prev = None
    forward = _ForwardPass(pro, cascade, zi, axis, _fwd_states)
    for fwd in forward:
        if prev is not None:
            # Look-ahead pass (numerical.py:397-399): only its final state is
            # used.  A stable cascade forgets where it started after `settle`
            # samples, so filtering the first `settle` samples of the next chunk
            # backwards from zi * (that sample) leaves the same state as the
            # reference's run from the chunk's far end, to ~1e-18 relative.
            look = cascade.look_ahead(zi, fwd)
            y = cascade.run(prev, look, reverse=True, out=_out)
            forward.release(prev)
            yield y
        prev = fwd
    if prev is not None and not _drop_last:
        last = cascade.state_from_sample(zi, prev, prev.shape[1] - 1)
        yield cascade.run(prev, last, reverse=True, out=_out)


class _ForwardPass:
    """The global forward pass of a forward-backward filter (numerical.py:394-396,
    :503-506), chunk by chunk, launched on its own CUDA stream: it depends only on the
    input and on the forward pass of the chunk before, so on the device it overlaps the
    backward pass of the previous chunk and whatever the consumer launches after it
    (the host issues chunk k+1's forward pass right after chunk k-1's backward pass).
    Iterating yields the forward output F of every chunk, ready for the current stream;
    ``release(F)`` hands its buffer back once the consumer has launched its last reader.

    F lives in a small pool of buffers owned by this object and allocated on the
    CONSUMER's stream; the side stream borrows them, fenced by the event recorded at
    ``release``.  (Blocks of the caching allocator passed between streams with
    record_stream are not reusable until the other stream has drained -- the allocator
    then grows with synchronous cudaMallocs; and an event recorded on the consumer's
    stream when a block is ALLOCATED would make the forward pass wait for everything
    queued there, which is exactly the work it is meant to overlap.)"""

    def __init__(self, pro, cascade, zi, axis, fwd_states=None):
        import os

        self.pro, self.cascade, self.zi, self.axis = pro, cascade, zi, axis
        self.states = fwd_states
        # Opt-in (OSZ_FWD_STREAM=1): measured on B200 at +1.4 % (256 rows) / +2 % (32 rows) --
        # both kernels are sized to fill the SMs, so they time-slice rather than overlap --
        # and the per-kernel CUDA-event times of concurrent kernels stop adding up to the
        # step, which is what bench.py's roofline accounting is built on.
        self.side = (dv.side_stream("iir-forward")
                     if os.environ.get("OSZ_FWD_STREAM", "0") == "1" else None)
        self.free = []                      # (buffer, event after its last reader)
        self.rows = _layout_of(pro, axis).rows

    def _take(self, n):
        for i, (buf, ev) in enumerate(self.free):
            if buf.shape[1] >= n:
                del self.free[i]
                return buf, ev
        return dv.empty_rows((self.rows, n)), dv.record_event()

    def release(self, fwd):
        buf = getattr(fwd, "_osz_buf", None)
        if buf is not None:
            self.free.append((buf, dv.record_event()))

    def __iter__(self):
        side, cascade = self.side, self.cascade
        chunks = iter(device_chunks(self.pro, self.axis))
        while True:
            with dv.on_stream(side):
                chunk = next(chunks, None)   # upstream stages / uploads run on the side stream too
            if chunk is None:
                return
            n = chunk.shape[1]
            buf, fence = self._take(n)
            view = buf if buf.shape[1] == n else buf[:, :n]
            with dv.on_stream(side):
                if side is not None and fence is not None:
                    side.wait_event(fence)
                if self.states is None:
                    self.states = cascade.state_from_sample(self.zi, chunk, 0)
                fwd = cascade.run(chunk, self.states, out=lambda r, m, view=view: view)
                done = dv.record_event() if side is not None else None
            if side is not None:
                dv.torch().cuda.current_stream().wait_event(done)
            fwd._osz_buf = buf
            yield fwd


def _sosfiltfilt_device(pro, sos, axis, _out=None, _free=False):
    dv.require_cuda()
    cascade = _Cascade(sos)
    zi = sps.sosfilt_zi(cascade.sos)
    yield from _filtfilt_device(pro, cascade, zi, axis, _out)


@_gpu_genfunc(_sosfiltfilt_device, _same_layout)
def sosfiltfilt(pro, sos, axis):
    """Forward-backward SOS filter with the reference's chunk-dependent
    semantics (numerical.py:338-411): the output depends on ``pro.chunksize``
    exactly as the reference's does."""


def _ba_order(coeffs):
    b, a = coeffs
    return int(max(len(np.atleast_1d(b)), len(np.atleast_1d(a))) - 1)


def _ba_to_sos(coeffs):
    """(b, a) of order <= 2 as one DF2T biquad -- scipy's lfilter recurrence for
    order 2 is the sosfilt section recurrence."""
    b, a = (np.atleast_1d(np.asarray(c, dtype=np.float64)) for c in coeffs)
    if max(len(b), len(a)) > 3:
        raise ValueError("_ba_to_sos: order above 2")
    b = np.pad(b, (0, 3 - len(b)))
    a = np.pad(a, (0, 3 - len(a)))
    return np.concatenate([b, a])[None, :]


class _TfFilter:
    """A (b, a) filter above second order behind the interface the shared
    drivers use of `_Cascade` (run / state_from_sample / zero_state / settle).
    The state of a row is scipy's lfilter `zi` vector (order,)."""

    def __init__(self, coeffs):
        b, a = (np.atleast_1d(np.asarray(c, dtype=np.float64)) for c in coeffs)
        self.plan = dv.TfPlan.cached(b, a)
        self.nstate = self.plan.nstate
        # DF2T zero-input transition (scipy lfilter's state): z_i' = z_(i+1) - a_(i+1) z_0
        k = self.nstate
        an = np.zeros(k + 1, dtype=np.longdouble)
        an[:len(a)] = np.asarray(a, dtype=np.longdouble) / np.longdouble(a[0])
        T = np.zeros((k, k), dtype=np.longdouble)
        for i in range(k):
            T[i, 0] = -an[i + 1]
            if i + 1 < k:
                T[i, i + 1] = 1
        self.settle = _settle_from_transition(T)

    def zero_state(self, rows):
        return dv.zeros((rows, self.nstate))

    def look_ahead(self, zi, fwd):
        """See _Cascade.look_ahead (state-only pass over the first `settle` samples)."""
        m = fwd.shape[1]
        if self.settle is not None and self.settle < m:
            m = self.settle
        look = self.state_from_sample(zi, fwd, m - 1)
        self.run(fwd[:, :m], look, reverse=True, want_output=False)
        return look

    def state_from_sample(self, zi, x, sample):
        return self.plan.state_from_sample(zi, x, sample)

    def run(self, x, state, reverse=False, want_output=True, out=None):
        dst = _new_rows(out, x.shape[0], x.shape[1]) if want_output else None
        return self.plan.run(x, state, reverse=reverse, want_output=want_output, out=dst)


def _tf_zi_rows(coeffs, zi, layout):
    """Reference lfilter zi layout (..., K-1, ...) -> (rows, K-1) state rows."""
    k = _ba_order(coeffs)
    zi = np.broadcast_to(np.asarray(zi, dtype=np.float64), layout.host_shape(k))
    zi = np.moveaxis(zi.reshape(layout.outer, k, layout.inner), 1, 2)
    return dv.from_host(np.ascontiguousarray(zi).reshape(layout.rows, k))


def _lfilter_zi_rows(coeffs, zi, layout):
    """Reference lfilter zi layout (..., K-1, ...) -> biquad state rows."""
    b, a = coeffs
    k = int(max(len(b), len(a)) - 1)
    zi = np.asarray(zi, dtype=np.float64)
    zi = np.broadcast_to(zi, layout.host_shape(k))
    pad = [(0, 0)] * zi.ndim
    pad[layout.axis] = (0, 2 - k)
    zi = np.pad(zi, pad)
    return _zi_to_rows(zi[None], layout, 1)


def _lfilter_device(pro, coeffs, axis, zi=None, _out=None, _free=False):
    dv.require_cuda()
    layout = _layout_of(pro, axis)
    if _ba_order(coeffs) > 2:
        filt = _TfFilter(coeffs)
        states = filt.zero_state(layout.rows) if zi is None else _tf_zi_rows(coeffs, zi, layout)
        for chunk in device_chunks(pro, axis, regrid=False):
            yield filt.run(chunk, states, out=_out)
        return
    cascade = _Cascade(_ba_to_sos(coeffs))
    if zi is None:
        states = cascade.zero_state(layout.rows)
    else:
        states = cascade.split_state(_lfilter_zi_rows(coeffs, zi, layout))
    for chunk in device_chunks(pro, axis, regrid=False):
        yield cascade.run(chunk, states, out=_out)


@_gpu_genfunc(_lfilter_device, _same_layout)
def lfilter(pro, coeffs, axis, zi=None):
    """Forward (b, a) filter with carried state (reference numerical.py:414-446).
    Second order and below (``Notch`` always is, filtering/iir.py:391) runs on
    the time-parallel biquad scan, higher orders on the sequential DF2T kernel."""


def _filtfilt_ba_device(pro, coeffs, axis, _out=None, _free=False):
    dv.require_cuda()
    z = np.atleast_1d(sps.lfilter_zi(*coeffs))        # numerical.py:487
    if _ba_order(coeffs) > 2:
        yield from _filtfilt_device(pro, _TfFilter(coeffs), z, axis, _out)
        return
    cascade = _Cascade(_ba_to_sos(coeffs))
    zi = np.zeros((1, 2))
    zi[0, :len(z)] = z
    yield from _filtfilt_device(pro, cascade, zi, axis, _out)


@_gpu_genfunc(_filtfilt_ba_device, _same_layout)
def filtfilt(pro, coeffs, axis):
    """Forward-backward (b, a) filter, chunk-dependent like the reference
    (numerical.py:449-520)."""


# ---------------------------------------------------------------------------
# polyphase resampling  (reference core/numerical.py:523-632)
# ---------------------------------------------------------------------------
def _resample_geometry(pro, L, M, axis):
    """Chunk size and yield boundaries of the reference (numerical.py:574-587,
    :617-632): csize <= N//3, multiple of M; the last two chunks are merged."""
    nsamp = pro.shape[axis]
    csize = int(pro.chunksize)
    if csize > nsamp // 3:
        csize = nsamp // 3
    if csize % M > 0:
        csize = int(np.ceil(csize / M) * M)
    nchunks = dv.ceil_div(nsamp, csize)
    return nsamp, csize, nchunks


def _resample_taps(L, M, fs, fir, kwargs):
    kwargs = dict(kwargs)
    cutoff = fs / (2 * max(L, M))
    fstop = kwargs.pop("fstop", cutoff + cutoff / 10)
    fpass = kwargs.pop("fpass", cutoff - cutoff / 10)
    gpass, gstop = kwargs.pop("gpass", 0.1), kwargs.pop("gstop", 40)
    return np.asarray(fir(fpass, fstop, fs, gpass, gstop).coeffs, dtype=np.float64)


class _Resampler:
    """Computes output samples [o_lo, o_hi) of ``resample_poly(x, L, M, window=h)``
    from a window of the input held on the device."""

    def __init__(self, pro, h, L, M, axis):
        self.source, self.nsamp_in = pro, pro.shape[axis]
        self.plan = dv.UpfirdnPlan.cached(h, L, M)
        self.reach = len(h)                        # taps of the filter the input sees
        self.center = (len(h) - 1) // 2

    def compute(self, window, w_first, o_lo, o_hi, out):
        return self.plan.run(window, w_first, o_lo, o_hi - o_lo, out=out)

    def truncate(self, nsamp):
        """The source ended after ``nsamp`` samples (short of its declared shape)."""
        self.nsamp_in = nsamp


class _FusedFirDecimator(_Resampler):
    """``downsample(FIR(x))`` as ONE decimating filter.

    An FIR stage in mode 'same' followed by decimation by M is, away from the
    recording's ends, a single FIR with taps c = h_fir * h_antialias evaluated
    at every M-th sample: K/M multiply-adds per input sample instead of a full
    rate FIR plus a decimator, and the full-rate intermediate (2 x 8 bytes per
    sample of HBM traffic) is never written.  The first and last few outputs
    see the 'same'-mode truncation of the intermediate signal; they are
    computed with the two unfused kernels on the edge samples only, so the
    result equals the two-stage result everywhere (to rounding)."""

    def __init__(self, fir_pro, fir_taps, h, M, axis):
        inner = fir_pro.data.args[0]
        self.source, self.nsamp_in = inner, inner.shape[axis]
        self.k1, self.k2, self.M = len(fir_taps), len(h), M
        self.left = (self.k1 - 1) // 2             # 'same' cuts of the FIR stage
        self.right = self.k1 - 1 - self.left
        self.half2 = (self.k2 - 1) // 2
        self.ny = self.nsamp_in                    # 'same': FIR output length
        fused = np.convolve(np.asarray(h, dtype=np.float64), fir_taps)
        self.plan = dv.UpfirdnPlan.cached(fused, 1, M)
        self.fir_plan = dv.FirPlan.cached(fir_taps)
        self.dec_plan = dv.UpfirdnPlan.cached(h, 1, M)
        self.reach = len(fused)
        self.center = (len(fused) - 1) // 2
        assert self.center == self.half2 + self.left
        # outputs j_lo .. j_hi see an untruncated intermediate signal
        self.j_lo = dv.ceil_div(self.k2 - 1 - self.half2, M)
        self.j_hi = (self.ny - 1 - self.half2) // M

    def truncate(self, nsamp):
        self.nsamp_in = self.ny = nsamp
        self.j_hi = (self.ny - 1 - self.half2) // self.M

    def _fir_span(self, window, w_first, u0, u1):
        """FIR 'same' output samples [u0, u1) from the input window."""
        rows = window.shape[0]
        lo = u0 + self.left - (self.k1 - 1)        # first input sample needed
        hi = u1 + self.left                        # exclusive
        zl, zr = max(0, -lo), max(0, hi - self.nsamp_in)
        core = window[:, max(lo, 0) - w_first:min(hi, self.nsamp_in) - w_first]
        buf = dv.cat_time([dv.zeros_rows(rows, zl) if zl else None, core,
                           dv.zeros_rows(rows, zr) if zr else None])
        return self.fir_plan.run(buf, u1 - u0)

    def left_edge(self, window, w_first, o_lo, e, out):
        """Outputs [o_lo, e), e <= j_lo: they see the 'same'-mode truncation of the
        intermediate signal at the recording's start -- two unfused kernels."""
        u1 = (e - 1) * self.M + self.half2 + 1
        y = self._fir_span(window, w_first, 0, u1)
        self.dec_plan.run(y, 0, o_lo, e - o_lo, out=out)

    def right_edge(self, window, w_first, s0, o_hi, out):
        """Outputs [s0, o_hi), s0 > j_hi: the truncation at the recording's end."""
        u0 = max(0, s0 * self.M + self.half2 - (self.k2 - 1))
        y = self._fir_span(window, w_first, u0, self.ny)
        self.dec_plan.run(y, u0, s0, o_hi - s0, out=out)

    def compute(self, window, w_first, o_lo, o_hi, out):
        if out is None:
            out = dv.empty_rows((window.shape[0], o_hi - o_lo))
        a, b = max(o_lo, self.j_lo), min(o_hi, self.j_hi + 1)
        if b > a:
            self.plan.run(window, w_first, a, b - a, out=out[:, a - o_lo:b - o_lo])
        if o_lo < self.j_lo:                       # left edge of the recording
            e = min(o_hi, self.j_lo)
            self.left_edge(window, w_first, o_lo, e, out[:, :e - o_lo])
        if o_hi > self.j_hi + 1:                   # right edge
            s0 = max(o_lo, self.j_hi + 1)
            self.right_edge(window, w_first, s0, o_hi, out[:, s0 - o_lo:])
        return out


def _fusable_fir(pro, L, M, ntaps2, axis):
    """The (fir taps) of ``pro`` when it is an openseize_b200 FIR stage that can
    be folded into the decimator that consumes it, else None."""
    import os

    if L != 1 or M < 2 or os.environ.get("OSZ_FUSE", "1") == "0":
        return None
    if not isinstance(pro, GenProducer) or pro.kwargs:
        return None
    func = pro.data
    if not isinstance(func, functools.partial) or func.func is not oaconvolve:
        return None
    args, kw = func.args, func.keywords or {}
    if len(args) != 4 or set(kw) - {"nfft_factor"}:
        return None
    inner, taps, fir_axis, mode = args
    taps = np.asarray(taps, dtype=np.float64)
    nd = len(pro.shape)
    if mode != "same" or normalize_axis(fir_axis, nd) != normalize_axis(axis, nd):
        return None
    if len(taps) % 2 == 0 or ntaps2 % 2 == 0:      # centres must add up exactly
        return None
    if not isinstance(inner, Producer) or inner.shape[axis] < 4 * (len(taps) + ntaps2 + M):
        return None
    return taps


# chunks that went through the fused IIR + FIR + decimator kernel / its unfused
# fallback (diagnostics for tests and bench.py)
FUSED_STATS = {"fused_chunks": 0, "fallback_chunks": 0}


def _fusable_iir(pro, axis):
    """(input producer, cascade, zi) when ``pro`` is an openseize_b200 forward-backward
    IIR stage -- ``sosfiltfilt``, or ``filtfilt`` of order <= 2 (``Notch``) -- whose
    last pass can run inside the decimator that consumes it, else None."""
    import os

    # Opt-in (OSZ_FUSE_IIR=1).  Measured on B200 (256 x 1e6, notch + 1231 taps, M = 25,
    # profiles/r02_ncu_summary.md): the fused kernel moves 8.3 bytes per sample instead of
    # 24.3, but the scan's ~55 instructions per sample and the decimator's MMA stream
    # share one SM's issue slots (39 % issue-active, the MMA group starved 20 % of the
    # time): 2.6 ms against 0.82 + 1.30 ms for the backward pass and the tensor-core
    # decimator run one after the other.
    if os.environ.get("OSZ_FUSE_IIR", "0") != "1" or dv.IO != "float64":
        return None
    if not isinstance(pro, GenProducer) or pro.kwargs:
        return None
    func = pro.data
    if not isinstance(func, functools.partial) or func.keywords:
        return None
    if func.func is not sosfiltfilt and func.func is not filtfilt:
        return None
    if len(func.args) != 3:
        return None
    inner, coeffs, iir_axis = func.args
    nd = len(pro.shape)
    if not isinstance(inner, Producer) or normalize_axis(iir_axis, nd) != normalize_axis(axis, nd):
        return None
    if func.func is sosfiltfilt:
        cascade = _Cascade(coeffs)
        zi = sps.sosfilt_zi(cascade.sos)
    else:
        if _ba_order(coeffs) > 2:
            return None
        z = np.atleast_1d(sps.lfilter_zi(*coeffs))        # numerical.py:487
        cascade = _Cascade(_ba_to_sos(coeffs))
        zi = np.zeros((1, 2))
        zi[0, :len(z)] = z
    if len(cascade.plans) != 1:
        return None
    return inner, cascade, zi


def _iir_fir_decimate_device(iir, fir_pro, fir_taps, h, M, axis, _out=None):
    """``downsample(FIR_same(IIR_dephase(x)))`` with the IIR's backward pass, the FIR
    and the decimator as ONE kernel per chunk: the forward pass
    writes F, the fused kernel reads it backwards, keeps the filtered samples in
    shared memory and emits only the decimated outputs -- the full-rate backward
    output (8 bytes written + 8 read per sample) never reaches HBM.

    Chunk semantics are the reference's (numerical.py:338-411,449-520): one global
    forward pass, each chunk's backward pass started from the state the look-ahead
    over the next forward chunk leaves (see _filtfilt_device).  Outputs whose window
    straddles two chunks (or two time spans of a chunk) are computed from the K-1
    edge samples every span exports; the recording's first and last few outputs, which
    see the 'same'-mode truncation, by the unfused kernels on those edge samples."""
    inner, cascade, zi = iir
    stage = _FusedFirDecimator(fir_pro, fir_taps, h, M, axis)
    sos_plan, ufd = cascade.plans[0], stage.plan
    rows = _layout_of(inner, axis).rows
    n_in, K, half = stage.nsamp_in, stage.reach, stage.center
    total_out = dv.ceil_div(n_in, M)
    st = {"done": 0, "tail": None}

    def emit(F, state, first, last):
        n = F.shape[1]
        done = st["done"]
        j_to = total_out if last else min(total_out, max(done, (first + n - 1 - half) // M + 1))
        nspan = dv.sosdec_spans(sos_plan, ufd, rows, n) if ufd.kernel == "mma" else 0
        FUSED_STATS["fused_chunks" if nspan else "fallback_chunks"] += 1
        if nspan == 0:
            # chunk too short for the fused tile (or no tensor-core geometry): the
            # backward pass in full, then the decimator on [previous tail | chunk]
            y = cascade.run(F, state, reverse=True)
            tail = st["tail"] if st["tail"] is not None else dv.zeros_rows(rows, K - 1)
            window = dv.cat_time([tail, y])
            st["tail"] = window[:, -(K - 1):]
            st["done"] = j_to
            if j_to > done:
                out = _new_rows(_out, rows, j_to - done)
                return stage.compute(window, first - (K - 1), done, j_to, out)
            return None
        out = _new_rows(_out, rows, max(j_to - done, 0))
        edges = dv.sosdec_exec(sos_plan, ufd, F, True, state[0], nspan, first, out, done)
        dv.sosdec_boundary(ufd, edges, True, st["tail"], last, n, first, out, done,
                           max(done, stage.j_lo), min(j_to - 1, stage.j_hi))
        if done < stage.j_lo and j_to > done:         # the recording's first outputs
            e = min(j_to, stage.j_lo)
            stage.left_edge(edges[:, nspan - 1, 0, :], first, done, e, out[:, :e - done])
        if j_to - 1 > stage.j_hi:                     # ... and its last ones
            s0 = max(done, stage.j_hi + 1)
            stage.right_edge(edges[:, 0, 1, :], first + n - (K - 1), s0, j_to, out[:, s0 - done:])
        st["tail"] = edges[:, 0, 1, :]                # span 0 of a backward pass is the last in time
        st["done"] = j_to
        return out if j_to > done else None

    prev, prev_first, pos = None, 0, 0
    forward = _ForwardPass(inner, cascade, zi, axis)
    for fwd in forward:
        if prev is not None:
            look = cascade.look_ahead(zi, fwd)        # see _filtfilt_device
            out = emit(prev, look, prev_first, False)
            forward.release(prev)
            if out is not None:
                yield out
        prev, prev_first = fwd, pos
        pos += fwd.shape[1]
    if prev is not None:
        stage.truncate(pos)                           # what arrived is the recording
        total_out = dv.ceil_div(pos, M)
        last = cascade.state_from_sample(zi, prev, prev.shape[1] - 1)
        out = emit(prev, last, prev_first, True)
        if out is not None:
            yield out


def _polyphase_device(pro, L, M, fs, fir, axis, _out=None, _free=False, **kwargs):
    dv.require_cuda()
    nsamp = pro.shape[axis]
    if M >= nsamp:
        raise ValueError("Decimation factor must M={} be < pro.shape[{}] = {}"
                         .format(M, axis, nsamp))
    nsamp, csize, nchunks = _resample_geometry(pro, L, M, axis)
    h = _resample_taps(L, M, fs, fir, kwargs)
    total_out = dv.ceil_div(nsamp * L, M)
    per_chunk = csize * L // M
    rows = _layout_of(pro, axis).rows

    src = producer(pro, csize, axis)               # same mutation as numerical.py:590
    # scipy.signal.resample_poly divides up and down by their gcd before it scales
    # the taps by `up` (the chunk grid above and the tap design keep the caller's
    # L and M, as the reference's do): resample by 2/4 applies h, not 2 * h[::2]
    g = int(np.gcd(int(L), int(M)))
    L, M = int(L) // g, int(M) // g
    fir_taps = _fusable_fir(src, L, M, len(h), axis)
    if fir_taps is not None and _free:
        iir = _fusable_iir(src.data.args[0], axis)
        if iir is not None:
            yield from _iir_fir_decimate_device(iir, src, fir_taps, h, M, axis, _out)
            return
    stage = (_FusedFirDecimator(src, fir_taps, h, M, axis) if fir_taps is not None
             else _Resampler(src, h, L, M, axis))
    n_in, center, reach = stage.nsamp_in, stage.center, stage.reach
    ring = _TimeRing(rows)                         # input window and its first global index
    w_first = 0
    emitted = 0                                    # yields done (the reference makes nchunks-1)
    done = 0                                       # outputs already yielded (free mode)
    seen = 0                                       # input samples received

    def trim(o_next):
        # oldest input sample output o_next touches (2 samples of slack)
        nonlocal w_first
        keep_from = max((o_next * M + center - (reach - 1)) // L - 2, w_first)
        ring.drop(keep_from - w_first)
        w_first = keep_from

    # The values do not depend on the input blocking (one global resample_poly,
    # SURVEY 8a5), only the yield boundaries do: take upstream blocks as they come.
    for chunk in device_chunks(stage.source, axis, regrid=False, alloc=ring):
        ring.push(chunk)
        seen += chunk.shape[1]
        if _free:
            # A device consumer takes blocks of any size: emit every output whose
            # input support has arrived, so the ring only ever holds the filter's
            # reach instead of waiting for the reference's chunk boundary.
            o_hi = total_out if seen >= n_in else min(total_out, max(
                done, ((seen - 1) * L - center) // M + 1))
            if o_hi > done:
                out = _new_rows(_out, rows, o_hi - done)
                yield stage.compute(ring.window(), w_first, done, o_hi, out)
                done = o_hi
                trim(o_hi)
            continue
        while emitted < nchunks - 1:
            last_yield = emitted == nchunks - 2
            o_lo = emitted * per_chunk
            o_hi = total_out if last_yield else (emitted + 1) * per_chunk
            # newest input sample output o_hi-1 touches
            need_hi = n_in if last_yield else min(n_in, ((o_hi - 1) * M + center) // L + 1)
            if seen < need_hi:
                break
            out = _new_rows(_out, rows, o_hi - o_lo)
            yield stage.compute(ring.window(), w_first, o_lo, o_hi, out)
            emitted += 1
            done = o_hi
            trim(o_hi)
    if 0 < seen < n_in:
        # the source ran out before its declared length: what arrived is the whole
        # recording (the reference's chunk loop ends with its iterators)
        stage.truncate(seen)
        total_out = dv.ceil_div(seen * L, M)
        if total_out > done:
            out = _new_rows(_out, rows, total_out - done)
            yield stage.compute(ring.window(), w_first, done, total_out, out)


def _polyphase_layout(pro, L, M, fs, fir, axis, **kwargs):
    return _layout_of(pro, axis)


@_gpu_genfunc(_polyphase_device, _polyphase_layout)
def polyphase_resample(pro, L, M, fs, fir, axis, **kwargs):
    """Resample by L/M with a Kaiser anti-alias filter (reference
    numerical.py:523-632).  Yields ``nchunks - 1`` arrays of ``csize*L/M``
    samples (the last one takes the remainder), bit-identical in length and
    index to the reference."""


# ---------------------------------------------------------------------------
# spectra  (reference core/numerical.py:635-1087)
# ---------------------------------------------------------------------------
def _spec_norm(window, fs, scaling):
    if scaling == "spectrum":
        return 1 / np.sum(window) ** 2
    if scaling == "density":
        return 1 / (fs * np.sum(window ** 2))
    raise ValueError("Unknown scaling: {}".format(scaling))


def _spec_plan(fs, nfft, window, overlap, detrend, scaling):
    nfft = int(nfft)
    stride = nfft - int(nfft * overlap)
    coeffs = sps.get_window(window, nfft)
    norm = _spec_norm(coeffs, fs, scaling)
    return dv.SpecPlan.cached(nfft, stride, coeffs, detrend, norm)


def _segment_batches(pro, axis, plan, pad_left=0, pad_right=0, batch_samples=1 << 24):
    """Yield (device rows, nseg): contiguous spans holding ``nseg`` whole
    windows starting at stride multiples -- the device twin of the FIFO walk
    in _spectra_estimatives (reference numerical.py:817-849).  Small chunks are
    batched so each launch carries enough windows to fill the GPU."""
    rows = _layout_of(pro, axis).rows
    ring = _TimeRing(rows, capacity=batch_samples // max(rows, 1))
    ring.push_zeros(pad_left)

    def flush():
        nseg = plan.nseg_available(ring.size)
        if nseg > 0:
            yield ring.window(), nseg
            ring.drop(nseg * plan.stride)

    for chunk in device_chunks(pro, axis, regrid=False, alloc=ring):
        ring.push(chunk)
        if ring.size * rows >= batch_samples:
            yield from flush()
    ring.push_zeros(pad_right)
    yield from flush()


def modified_dft(arr, fs, nfft, window, axis, detrend, scaling):
    """Windowed DFT of one in-memory segment (reference numerical.py:635-718).
    Returns (freqs, X) with X complex128 of length nfft//2+1 along ``axis``."""
    arr = np.asarray(arr, dtype=np.float64)
    axis = normalize_axis(axis, arr.ndim)
    nfft = int(nfft)
    nsamp = arr.shape[axis]
    if nfft < nsamp:
        arr = arr[..., :nfft] if axis == arr.ndim - 1 else np.take(arr, np.arange(nfft), axis)
    return _single_segment(arr, fs, nfft, window, axis, detrend, scaling, complex_=True)


def periodogram(arr, fs, nfft=None, window="hann", axis=-1, detrend="constant",
                scaling="density"):
    """Power spectrum of one in-memory segment (reference numerical.py:721-796)."""
    arr = np.asarray(arr, dtype=np.float64)
    axis = normalize_axis(axis, arr.ndim)
    nfft = arr.shape[axis] if not nfft else int(nfft)
    if nfft < arr.shape[axis]:
        arr = arr[..., :nfft] if axis == arr.ndim - 1 else np.take(arr, np.arange(nfft), axis)
    return _single_segment(arr, fs, nfft, window, axis, detrend, scaling, complex_=False)


def _single_segment(arr, fs, nfft, window, axis, detrend, scaling, complex_):
    nsamp = arr.shape[axis]
    if nsamp == 0:
        # the reference's detrend of an empty slice warns (numpy "Mean of empty
        # slice") before np.fft raises (tests/test_spectra.py:140-165): same behaviour
        import warnings

        warnings.warn("Mean of empty slice.", RuntimeWarning, stacklevel=3)
        raise ValueError("cannot estimate the spectrum of an empty array")
    if nsamp < nfft:
        # zero-padded transform (reference numerical.py:688-699): detrend and
        # window the nsamp samples, pad with zeros, transform with a unit window
        layout = dv.Layout(arr.shape, axis)
        coeffs = sps.get_window(window, nsamp)
        padded = dv.spec_prepare(dv.upload(arr, layout), nsamp, nfft, coeffs, detrend)
        plan = dv.SpecPlan.cached(nfft, nfft, np.ones(nfft), None,
                                  _spec_norm(coeffs, fs, scaling))
        out = plan.segments(padded, 1, complex_)[0]
        res = dv.download(out, layout, complex_).get()
        return np.fft.rfftfreq(nfft, d=1 / fs), np.array(res)
    layout = dv.Layout(arr.shape, axis)
    coeffs = sps.get_window(window, nsamp)
    plan = dv.SpecPlan.cached(nfft, nfft, coeffs, detrend, _spec_norm(coeffs, fs, scaling))
    rows = dv.upload(arr, layout)
    out = plan.segments(rows, 1, complex_)[0]
    res = dv.download(out, layout, complex_).get()
    return np.fft.rfftfreq(nfft, d=1 / fs), np.array(res)


def _estimatives_device(pro, fs, nfft, window, overlap, axis, detrend, scaling, func,
                        pad_left=0, pad_right=0, _out=None, _free=False, **kwargs):
    """Device rows (rows, nfreq[, 2]) per window, in order."""
    complex_ = func is modified_dft
    if func is not modified_dft and func is not periodogram:
        raise TypeError("func must be numerical.periodogram or numerical.modified_dft")
    for out, nseg in _estimatives_batches(pro, fs, nfft, window, overlap, axis, detrend, scaling,
                                          complex_, pad_left, pad_right):
        for k in range(nseg):
            yield out[k]


def _estimatives_batches(pro, fs, nfft, window, overlap, axis, detrend, scaling, complex_,
                         pad_left=0, pad_right=0):
    """(device tensor (nseg, rows, nfreq[, 2]), nseg) per batch of windows."""
    plan = _spec_plan(fs, nfft, window, overlap, detrend, scaling)
    for buf, nseg in _segment_batches(pro, axis, plan, pad_left, pad_right):
        yield plan.segments(buf, nseg, complex_), nseg


def _spectra_estimatives(pro, fs, nfft, window, overlap, axis, detrend, scaling, func,
                         pad_left=0, pad_right=0, **kwargs):
    """One estimate per nfft window, windows ``stride`` apart, trailing partial
    window dropped (reference numerical.py:799-849).  ``pad_left`` /
    ``pad_right`` are zeros put around the recording (STFT boundary/padded)."""
    complex_ = func is modified_dft
    layout = _layout_of(pro, axis)
    if layout.inner == 1 and (func is modified_dft or func is periodogram):
        # sample axis last: a batch of windows is one contiguous device block
        # (nseg, rows, nfreq[, 2]) -- ONE device-to-host copy per batch instead of one
        # per window (a thousand 1 MB copies cost more in launches than in PCIe time),
        # handed to the consumer one batch behind the kernels
        pending = None
        for out, nseg in _estimatives_batches(pro, fs, nfft, window, overlap, axis, detrend,
                                              scaling, complex_, pad_left, pad_right):
            nxt = (dv.download_block(out), nseg)
            if pending is not None:
                yield from _split_windows(pending, layout, complex_)
            pending = nxt
        if pending is not None:
            yield from _split_windows(pending, layout, complex_)
        return
    gen = _estimatives_device(pro, fs, nfft, window, overlap, axis, detrend, scaling, func,
                              pad_left, pad_right)
    yield from _to_host(gen, layout, complex_)


def _split_windows(pending, layout, complex_):
    block, nseg = pending
    host = block.get()                      # (nseg, rows, nfreq[, 2]) float64, pinned
    if complex_:
        host = host.view(np.complex128)[..., 0]
    if dv.IO == "float32":
        host = host.astype(np.complex64 if complex_ else np.float32)
    nfreq = host.shape[2]
    for k in range(nseg):
        yield host[k].reshape(layout.host_shape(nfreq))


_spectra_estimatives.device = _estimatives_device


def welch(pro, fs, nfft, window, overlap, axis, detrend, scaling):
    """(freqs, producer of one periodogram per Welch segment) -- reference
    numerical.py:852-947, including its quirk that the producer's shape carries
    the number of segments along ``axis``."""
    genfunc = functools.partial(_spectra_estimatives, pro, fs, nfft, window, overlap, axis,
                                detrend, scaling, func=periodogram)
    freqs = np.fft.rfftfreq(nfft, 1 / fs)
    nsegs = int((pro.shape[axis] - nfft) // (nfft * (1 - overlap)) + 1)
    shape = list(pro.shape)
    shape[axis] = nsegs
    return freqs, producer(genfunc, chunksize=len(freqs), axis=axis, shape=shape)


def welch_sum(pro, fs, nfft, window, overlap, axis, detrend, scaling):
    """(segment count, DEVICE (rows, nfft//2+1) SUM of the segment
    periodograms).  The per-segment periodograms never leave the SM: |FFT|^2
    is accumulated on chip and only the sum is written.  This is the quantity
    time-sharded ranks all-reduce (openseize_b200.sharding)."""
    dv.require_cuda()
    plan = _spec_plan(fs, nfft, window, overlap, detrend, scaling)
    layout = _layout_of(pro, axis)
    psd_sum = dv.zeros((layout.rows, plan.nfreq))
    cnt = 0
    for buf, nseg in _segment_batches(pro, axis, plan):
        plan.welch_accum(buf, nseg, psd_sum)
        cnt += nseg
    return cnt, psd_sum


def welch_mean(pro, fs, nfft, window, overlap, axis, detrend, scaling):
    """Fused Welch estimate: (segment count, mean periodogram ndarray).  Equals
    the reference's running mean over ``welch``'s producer
    (spectra/estimators.py:150-152) up to rounding."""
    cnt, psd_sum = welch_sum(pro, fs, nfft, window, overlap, axis, detrend, scaling)
    if cnt == 0:
        raise ValueError("psd: the data holds no complete nfft={} segment".format(nfft))
    layout = _layout_of(pro, axis)
    est = np.array(dv.download(psd_sum, layout).get()) / cnt
    return cnt, est.astype(np.float32) if dv.IO == "float32" else est


def stft(pro, fs, nfft, window, overlap, axis, detrend, scaling, boundary, padded):
    """(freqs, time, producer of one modified DFT per segment) -- reference
    numerical.py:950-1087.  Zero extension (``boundary``: nfft//2 both ends;
    ``padded``: one stride at the end when N % stride) happens on the device."""
    noverlap = int(nfft * overlap)
    stride = nfft - noverlap
    nsamp = pro.shape[axis]
    pad_left = nfft // 2 if boundary else 0
    pad_right = pad_left + ((stride if nsamp % stride else 0) if padded else 0)
    total = nsamp + pad_left + pad_right
    genfunc = functools.partial(_spectra_estimatives, pro, fs, nfft, window, overlap, axis,
                                detrend, scaling, func=modified_dft, pad_left=pad_left,
                                pad_right=pad_right)
    freqs = np.fft.rfftfreq(nfft, 1 / fs)
    nsegs = int((total - nfft) // (nfft * (1 - overlap)) + 1)
    shape = list(pro.shape)
    shape[axis] = nsegs
    if boundary:
        time = 1 / fs * np.arange(0, total - nfft + 1, nfft - noverlap)
    else:
        time = 1 / fs * np.arange(nfft // 2, total + 1 - nfft // 2, nfft - noverlap)
    return freqs, time, producer(genfunc, chunksize=len(freqs), axis=axis, shape=shape)
