"""Spreading one recording over the GPUs of a box: one process per GPU
(``torch.distributed``), work partitioned by channel, or by time for
recordings with few channels.

The reference has no distributed runtime.  What makes the hot
path shardable is that every operator is independent per 1-D slice along the
sample axis and needs only a bounded carry along time:

* channel sharding -- each rank runs the unchanged operator chain on its block
  of rows; no data-path communication at all;
* time sharding of the Welch PSD -- each rank owns a contiguous run of whole
  segments and all-reduces its partial periodogram SUM and segment count (the
  only collective on the path: NCCL over NVLink, a (rows, nfft//2+1) float64
  message -- 4 MB for 256 x 2049 -- so it is latency bound, not bandwidth
  bound);
* time sharding of FIR filtering and resampling -- each rank reads its span of
  the recording plus a filter-length halo straight from the host array (no
  GPU-to-GPU traffic) and keeps the outputs of its own span;
* time sharding of the IIR filters -- the recurrence carries state through
  time, so every rank first reduces its span to the state it would leave when
  entered at rest, the (rows, nsec, 2) summaries are all-gathered (kilobytes)
  and composed with the cascade's zero-input transition matrix, and each rank
  then filters its span from its true entering state.  The forward-backward
  filters additionally borrow the next span's first chunk, because the
  reference's backward pass looks one chunk ahead (numerical.py:394-403).

Everything here is host-side planning plus small collectives; the arithmetic
is the same GPU kernels.
"""

import numpy as np

from openseize_b200.core import device as dv
from openseize_b200.core import numerical as nm
from openseize_b200.core.arraytools import normalize_axis, slice_along_axis
from openseize_b200.core.producer import ArrayProducer, Producer, producer


def _dist():
    import torch.distributed as dist

    return dist


def world(group=None):
    """(rank, world_size) of this process; (0, 1) outside torch.distributed."""
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def split_range(n, parts):
    """``parts`` contiguous, near-equal [start, stop) ranges covering range(n);
    the first ``n % parts`` ranges are one longer.  Ranges may be empty."""
    base, extra = divmod(int(n), int(parts))
    bounds, start = [], 0
    for r in range(parts):
        stop = start + base + (1 if r < extra else 0)
        bounds.append((start, stop))
        start = stop
    return bounds


def channel_block(shape, axis, rank, size):
    """Slices selecting this rank's block of channels: the largest non-sample
    axis is split.  Returns (tuple of slices, split axis)."""
    axis = normalize_axis(axis, len(shape))
    others = [a for a in range(len(shape)) if a != axis]
    if not others:
        raise ValueError("a 1-D recording has no channel axis to shard; shard it in time")
    split = max(others, key=lambda a: shape[a])
    lo, hi = split_range(shape[split], size)[rank]
    index = [slice(None)] * len(shape)
    index[split] = slice(lo, hi)
    return tuple(index), split


def _block_chunks(pro, index):
    """Generating function of a channel block of any producer: every yielded chunk
    sliced to the block (module level so that the sharded producer pickles)."""
    for chunk in pro:
        yield chunk[index]


def shard_channels(data, chunksize, axis=-1, group=None):
    """Producer over this rank's block of channels.  Running any operator chain on
    it and concatenating the ranks' results along the split axis equals the
    single-process result (every operator is independent per 1-D slice along the
    sample axis: reference filtering/bases.py:169-172).

    * ndarray / array producer: a view of the block -- no copy;
    * reader producer (core/producer.py:213-264) whose reader has a ``channels``
      attribute (the EDF readers): a copy of the reader restricted to the block, so
      only this rank's channels are read and decoded -- the out-of-core case;
    * any other producer (generating functions, masked): the same chunks sliced to
      the block as they are produced."""
    import copy
    import functools

    from openseize_b200.core.producer import ReaderProducer

    rank, size = world(group)
    if isinstance(data, ArrayProducer):
        data = data.data
    if not isinstance(data, Producer):
        index, _ = channel_block(np.shape(data), axis, rank, size)
        return producer(np.asarray(data)[index], chunksize, axis)
    axis_n = normalize_axis(axis, len(data.shape))
    index, split = channel_block(data.shape, axis_n, rank, size)
    lo, hi, _ = index[split].indices(data.shape[split])
    shape = list(data.shape)
    shape[split] = hi - lo
    reader = getattr(data, "data", None)
    chans = getattr(reader, "channels", None)
    if isinstance(data, ReaderProducer) and chans is not None and len(data.shape) == 2 \
            and split == 0 and len(chans) == data.shape[0]:
        mine = copy.copy(reader)
        mine.channels = list(chans)[lo:hi]
        return producer(mine, chunksize, axis, **dict(data.kwargs))
    src = producer(data, data.chunksize, data.axis)         # (same object: chunk size kept)
    return producer(functools.partial(_block_chunks, src, index), chunksize, axis,
                    shape=tuple(shape))


def welch_segments(nsamples, nfft, overlap):
    """(number of whole Welch segments, stride) -- reference numerical.py:817-818."""
    stride = nfft - int(nfft * overlap)
    nseg = (nsamples - nfft) // stride + 1 if nsamples >= nfft else 0
    return nseg, stride


def time_span(nsamples, nfft, overlap, rank, size):
    """Samples [start, stop) holding exactly this rank's run of whole segments
    (consecutive spans overlap by nfft - stride); (0, 0) for a rank with none."""
    nseg, stride = welch_segments(nsamples, nfft, overlap)
    k0, k1 = split_range(nseg, size)[rank]
    if k1 <= k0:
        return 0, 0
    return k0 * stride, (k1 - 1) * stride + nfft


def psd_time_sharded(data, fs, axis=-1, resolution=0.5, window="hann", overlap=0.5,
                     detrend="constant", scaling="density", group=None):
    """Welch PSD of one recording with its TIME axis split over the ranks of
    ``group``.  Same signature and return value as ``spectra.estimators.psd``
    ((segment count, frequencies, estimate)); every rank gets the full result.

    Each rank sums the periodograms of its own segments on its GPU; one
    ``all_reduce(SUM)`` of that (rows, nfft//2+1) array plus the count combines
    them; the mean is taken after the reduction, so the estimate equals the
    single-process one up to the order of the floating-point sum."""
    t = dv.require_cuda()
    dist = _dist()
    data = _Source(data, axis)
    axis = data.axis
    rank, size = world(group)
    nfft = int(fs / resolution)
    nsamples = data.shape[axis]
    start, stop = time_span(nsamples, nfft, overlap, rank, size)
    layout = dv.Layout(data.shape, axis)
    if stop > start:
        cnt, psd_sum = nm.welch_sum(data.window(start, stop, int(fs)), fs, nfft, window, overlap,
                                    axis, detrend, scaling)
    else:
        cnt, psd_sum = 0, dv.zeros((layout.rows, nfft // 2 + 1))
    packed = t.cat([psd_sum.reshape(-1),
                    t.tensor([float(cnt)], dtype=t.float64, device=psd_sum.device)])
    if size > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    total = int(round(float(packed[-1].item())))
    expected, _ = welch_segments(nsamples, nfft, overlap)
    assert total == expected, (total, expected)
    if total == 0:
        raise ValueError("psd: the data holds no complete nfft={} segment".format(nfft))
    summed = packed[:-1].reshape(layout.rows, nfft // 2 + 1)
    estimate = np.array(dv.download(summed, layout).get()) / total
    return total, np.fft.rfftfreq(nfft, 1 / fs), estimate


# ---------------------------------------------------------------------------
# time sharding of the filters (few-channel recordings)
# ---------------------------------------------------------------------------
def _window_chunks(pro, a, b, axis):
    """Generating function of samples [a, b) of any producer: chunks outside the window
    are skipped as they are produced (a generating function cannot seek)."""
    pos = 0
    for chunk in pro:
        n = chunk.shape[axis]
        lo, hi = max(a - pos, 0), min(b - pos, n)
        if hi > lo:
            yield slice_along_axis(chunk, lo, hi, axis=axis)
        pos += n
        if pos >= b:
            return


class _Source:
    """A recording that can hand out time windows: an ndarray or array producer (views),
    a reader producer (``read(start, stop)`` of the window only -- the out-of-core
    case), or any other producer (its chunks, cut to the window as they come)."""

    def __init__(self, data, axis):
        from openseize_b200.core.producer import ReaderProducer

        if isinstance(data, ArrayProducer):
            data = data.data
        self.data = data
        self.kind = ("array" if not isinstance(data, Producer)
                     else "reader" if isinstance(data, ReaderProducer) else "producer")
        if self.kind == "array":
            self.data = np.asarray(data)
        self.shape = tuple(self.data.shape)
        self.ndim = len(self.shape)
        self.axis = normalize_axis(axis, self.ndim)

    def window(self, a, b, chunksize):
        """Producer of samples [a, b) along the sample axis."""
        import functools

        a, b = int(a), int(b)
        if self.kind == "array":
            return producer(slice_along_axis(self.data, a, b, axis=self.axis), chunksize, self.axis)
        if self.kind == "reader":
            pro = self.data
            return producer(pro.data, chunksize, self.axis, start=pro.start + a,
                            stop=pro.start + b, **dict(pro.kwargs))
        shape = list(self.shape)
        shape[self.axis] = b - a
        src = producer(self.data, self.data.chunksize, self.data.axis)
        return producer(functools.partial(_window_chunks, src, a, b, self.axis), chunksize,
                        self.axis, shape=tuple(shape))

    def array(self, a, b):
        """Samples [a, b) as an ndarray (small windows: first sample, halos)."""
        if self.kind == "array":
            return slice_along_axis(self.data, int(a), int(b), axis=self.axis)
        return self.window(a, b, max(int(b) - int(a), 1)).to_array()


def time_spans(n, size, align=1):
    """``size`` contiguous spans covering range(n) whose interior boundaries are
    multiples of ``align`` (the chunk grid for the chunk-dependent filters)."""
    units = -(-int(n) // int(align))
    return [(min(a * align, n), min(b * align, n)) for a, b in split_range(units, size)]


def _collective_device(group=None):
    """Where tensors of a collective have to live: the GPU under NCCL, the host
    under gloo (the CPU test harness)."""
    dist = _dist()
    return "cuda" if dist.get_backend(group) == "nccl" else "cpu"


def all_gather_arrays(local, group=None):
    """Every rank's float64 ndarray (or None) of a common trailing shape, gathered
    with tensor collectives -- one ``all_gather`` of the leading lengths, one of the
    arrays padded to the longest -- instead of pickled objects.  Returns the list of
    arrays by rank (None for ranks that contributed None)."""
    t = dv.torch()
    dist = _dist()
    rank, size = world(group)
    if size == 1:
        return [local]
    dev = _collective_device(group)
    # trailing shape: agreed through a MAX reduction (ranks with None send zeros)
    ndim_t = t.tensor([0 if local is None else local.ndim], dtype=t.int64, device=dev)
    dist.all_reduce(ndim_t, op=dist.ReduceOp.MAX, group=group)
    ndim = int(ndim_t.item())
    meta = t.zeros(1 + max(ndim, 1), dtype=t.int64, device=dev)
    if local is not None:
        meta[0] = 1
        meta[1:1 + local.ndim] = t.tensor(local.shape, dtype=t.int64)
    metas = [t.zeros_like(meta) for _ in range(size)]
    dist.all_gather(metas, meta, group=group)
    metas = [m.cpu().numpy() for m in metas]
    trail = None
    for m in metas:
        if m[0]:
            trail = tuple(int(v) for v in m[2:1 + ndim])
    lead = [int(m[1]) if m[0] else 0 for m in metas]
    if trail is None:
        return [None] * size
    width = int(np.prod(trail, dtype=np.int64)) if trail else 1
    longest = max(max(lead), 1)
    buf = t.zeros((longest, width), dtype=t.float64, device=dev)
    if local is not None and local.size:
        buf[:local.shape[0]] = t.from_numpy(
            np.ascontiguousarray(local, dtype=np.float64).reshape(local.shape[0], width)).to(dev)
    bufs = [t.empty_like(buf) for _ in range(size)]
    dist.all_gather(bufs, buf, group=group)
    out = []
    for m, n, b in zip(metas, lead, bufs):
        out.append(b[:n].cpu().numpy().reshape((n,) + trail) if m[0] else None)
    return out


def gather_time(local, axis, group=None):
    """Concatenate every rank's span along ``axis`` (every rank gets the whole
    result).  ``local`` may be None for a rank with an empty span.  Tensor
    collectives over the group's backend (NCCL: device tensors over NVLink)."""
    rank, size = world(group)
    if size == 1:
        return local
    lead = None if local is None else np.moveaxis(local, axis, 0)
    parts = [p for p in all_gather_arrays(lead, group) if p is not None]
    return np.moveaxis(np.concatenate(parts, axis=0), 0, axis)


def fir_time_sharded(data, window, chunksize, axis=-1, mode="same", group=None, gather=True):
    """``nm.oaconvolve`` of one recording with the OUTPUT's time axis split over
    the ranks.  Rank r convolves samples [o0 - (K-1), o1) of the input (zeros
    outside the recording, as numpy's convolution modes imply) and keeps outputs
    [o0, o1); the halo comes from the host array.  Returns the whole result
    (``gather``) or ``((o0, o1), this rank's span)``."""
    data = _Source(data, axis)
    axis = data.axis
    window = np.asarray(window, dtype=np.float64)
    k, n = len(window), data.shape[axis]
    if n < k:
        raise ValueError("oaconvolve: data length {} along axis is shorter than the {} taps"
                         .format(n, k))
    left, right = nm._mode_cuts(k, mode)
    total = n + k - 1 - left - right                  # output samples of this mode
    rank, size = world(group)
    o0, o1 = split_range(total, size)[rank]
    local = None
    if o1 > o0:
        f0, f1 = left + o0, left + o1                 # full-convolution indices wanted
        in_lo, in_hi = max(f0 - (k - 1), 0), min(f1, n)
        if in_hi - in_lo < k:                         # a sliver: widen the halo to K samples
            in_lo = max(in_hi - k, 0)
            in_hi = min(in_lo + k, n)
        blocks = list(nm.oaconvolve(data.window(in_lo, in_hi, chunksize), window, axis, "full"))
        full = np.concatenate(blocks, axis=axis)      # full[j] is full-convolution index in_lo + j
        local = slice_along_axis(full, f0 - in_lo, f1 - in_lo, axis=axis)
    if not gather:
        return (o0, o1), local
    return gather_time(local, axis, group)


def resample_time_sharded(data, L, M, fs, chunksize, axis=-1, group=None, gather=True, **kwargs):
    """``nm.polyphase_resample`` with the output's time axis split over the
    ranks.  The reference's chunked resampler equals one global
    ``scipy.signal.resample_poly`` call for every chunking, so a
    rank resamples its input span widened by the filter reach on both sides,
    aligned to the decimation grid, and keeps exactly its outputs."""
    from math import gcd

    from openseize_b200.filtering.fir import Kaiser

    data = _Source(data, axis)
    axis = data.axis
    n = data.shape[axis]
    g = gcd(int(L), int(M))
    up, down = int(L) // g, int(M) // g
    total = -(-n * up // down)                        # ceil(N L / M), resampling.py:90-92
    h = nm._resample_taps(L, M, fs, Kaiser, kwargs)
    reach = -(-(len(h) - 1) // (2 * up)) + 1           # input samples a tap can reach either side
    rank, size = world(group)
    o0, o1 = split_range(total, size)[rank]
    local = None
    if o1 > o0:
        in_lo = max((o0 * down) // up - reach, 0)
        in_lo -= in_lo % down                         # output index in_lo*up/down is an integer
        in_hi = min(-(-(o1 * down) // up) + reach + down, n)
        # the sub-recording must satisfy the reference's own size rules (numerical.py:569-576)
        in_hi = min(max(in_hi, in_lo + 3 * down + len(h)), n)
        in_lo = max(min(in_lo, in_hi - 3 * down - len(h)), 0)
        in_lo -= in_lo % down
        sub_pro = data.window(in_lo, in_hi, chunksize)
        blocks = list(nm.polyphase_resample(sub_pro, L, M, fs, Kaiser, axis, **kwargs))
        y = np.concatenate(blocks, axis=axis)
        first = in_lo * up // down                    # global index of y's first sample
        local = slice_along_axis(y, o0 - first, o1 - first, axis=axis)
        assert local.shape[axis] == o1 - o0, (local.shape, o0, o1, first)
    if not gather:
        return (o0, o1), local
    return gather_time(local, axis, group)


def cascade_transition(sos):
    """One-step zero-input state transition of a DF2T biquad cascade: the
    (2 nsec, 2 nsec) matrix T with state' = T state when the input sample is 0.
    State order is scipy's: (z0, z1) of section 0, then section 1, ...  Derived
    from the recurrence of SURVEY.md 8a2 (y = b0 x + z0; z0' = b1 x - a1 y + z1;
    z1' = b2 x - a2 y; the next section's x is this section's y)."""
    sos = np.atleast_2d(np.asarray(sos, dtype=np.longdouble))
    sos = sos / sos[:, 3:4]
    nsec = sos.shape[0]
    T = np.zeros((2 * nsec, 2 * nsec), dtype=np.longdouble)
    for j in range(2 * nsec):
        state = np.zeros((nsec, 2), dtype=np.longdouble)
        state.reshape(-1)[j] = 1
        x = np.longdouble(0)
        nxt = np.zeros_like(state)
        for s in range(nsec):
            b0, b1, b2, _, a1, a2 = sos[s]
            y = b0 * x + state[s, 0]
            nxt[s, 0] = b1 * x - a1 * y + state[s, 1]
            nxt[s, 1] = b2 * x - a2 * y
            x = y
        T[:, j] = nxt.reshape(-1)
    return T


def _matrix_power(T, n):
    out = np.eye(T.shape[0], dtype=T.dtype)
    base = T.copy()
    n = int(n)
    while n:
        if n & 1:
            out = base @ out
        base = base @ base
        n >>= 1
    return out


def _states_to_host(states):
    t = dv.torch()
    return t.cat([s for s in states], dim=1).cpu().numpy().astype(np.float64)


def _entering_states(cascade, data, spans, rank, axis, chunksize, first_state, group):
    """Forward state entering this rank's span.  Every rank filters its own span
    from rest and keeps only the final state f_r; with Phi_r = T^(span length),
    s_0 = first_state and s_(r+1) = Phi_r s_r + f_r  (superposition of the
    zero-state and zero-input responses)."""
    layout = dv.Layout(data.shape, axis)
    a, b = spans[rank]
    size = len(spans)
    f = np.zeros((layout.rows, cascade.nsec, 2))
    if size > 1 and b > a and rank < size - 1:
        states = cascade.zero_state(layout.rows)
        for chunk in nm.device_chunks(data.window(a, b, chunksize), axis, regrid=False):
            cascade.run(chunk, states, want_output=False)
        f = _states_to_host(states)
    finals = all_gather_arrays(f, group) if size > 1 else [f]
    T = cascade_transition(cascade.sos)
    s = np.asarray(first_state, dtype=np.longdouble).reshape(layout.rows, -1)
    for q in range(rank):
        qa, qb = spans[q]
        if qb > qa:
            phi = _matrix_power(T, qb - qa)
            s = s @ phi.T + finals[q].reshape(layout.rows, -1)
    return np.ascontiguousarray(s.astype(np.float64)).reshape(layout.rows, cascade.nsec, 2)


def iir_time_sharded(data, coeffs, chunksize, axis=-1, dephase=True, zi=None, fmt="sos",
                     group=None, gather=True):
    """``IIR.__call__`` semantics (filtering/bases.py:153-213) for one recording
    with its time axis split over the ranks on the chunk grid of ``chunksize``:
    ``nm.sosfiltfilt`` / ``nm.filtfilt`` when ``dephase`` else ``nm.sosfilt`` /
    ``nm.lfilter`` (``fmt`` 'sos' or 'ba'; (b, a) of second order at most).
    Results equal the single-process ones to rounding: the forward-backward
    filters keep the reference's dependence on the chunk grid."""
    import scipy.signal as sps

    data = _Source(data, axis)
    axis = data.axis
    n = data.shape[axis]
    chunksize = int(chunksize)
    layout = dv.Layout(data.shape, axis)
    if fmt == "sos":
        sos = np.atleast_2d(np.asarray(coeffs, dtype=np.float64))
        zi_ss = sps.sosfilt_zi(sos)
    else:
        sos = nm._ba_to_sos(coeffs)
        z = np.atleast_1d(sps.lfilter_zi(*coeffs))
        zi_ss = np.zeros((1, 2))
        zi_ss[0, :len(z)] = z
    cascade = nm._Cascade(sos)
    rank, size = world(group)
    spans = time_spans(n, size, chunksize)
    a, b = spans[rank]

    # state entering the recording
    if dephase:
        x0 = np.moveaxis(np.asarray(data.array(0, 1), dtype=np.float64).reshape(
            layout.outer, 1, layout.inner), 1, 2).reshape(layout.rows)
        first = zi_ss[None, :, :] * x0[:, None, None]            # zi * x0, numerical.py:385
    elif zi is None:
        first = np.zeros((layout.rows, cascade.nsec, 2))
    elif fmt == "sos":
        first = nm._zi_to_rows(zi, layout, cascade.nsec).cpu().numpy()
    else:
        first = nm._lfilter_zi_rows(coeffs, zi, layout).cpu().numpy()
    entering = _entering_states(cascade, data, spans, rank, axis, chunksize, first, group)

    local = None
    if b > a:
        states = cascade.split_state(dv.from_host(entering))
        if dephase:
            stop = min(b + chunksize, n)                          # borrow the look-ahead chunk
            gen = nm._filtfilt_device(data.window(a, stop, chunksize), cascade, zi_ss, axis,
                                      _fwd_states=states, _drop_last=stop > b)
        else:
            gen = (cascade.run(chunk, states)
                   for chunk in nm.device_chunks(data.window(a, b, chunksize), axis,
                                                 regrid=False))
        local = np.concatenate(list(nm._to_host(gen, layout)), axis=axis)
        assert local.shape[axis] == b - a
    if not gather:
        return (a, b), local
    return gather_time(local, axis, group)
