"""numpy/scipy restatement of openseize's chunked DSP algorithms.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Every function takes the *whole* signal as an ndarray plus the chunking
parameters the reference would have received through its producer
(``chunksize``, ``axis``) and returns the list of arrays the reference's
generator would have yielded, in order.  Concatenating that list along
``axis`` gives what ``Producer.to_array()`` returns.  All citations are
``/root/reference/src/openseize/...`` file:line.
"""

import math

import numpy as np
import scipy.signal as sps


# --------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------
def _ax(a, start=None, stop=None, axis=-1):
    """Slice along one axis (core/arraytools.py:43-58)."""
    sl = [slice(None)] * a.ndim
    sl[axis] = slice(start, stop)
    return a[tuple(sl)]


def split_chunks(x, chunksize, axis=-1):
    """Chunk grid of an ArrayProducer (core/producer.py:289-295): consecutive
    views of ``chunksize`` samples along ``axis``; the last may be shorter."""
    x = np.asarray(x)
    n = x.shape[axis]
    cs = int(chunksize)
    return [_ax(x, s, min(s + cs, n), axis) for s in range(0, n, cs)]


class _Fifo:
    """Concat-on-put / split-on-get queue (core/queues.py:9-70)."""

    def __init__(self, size, axis):
        self.size, self.axis, self.q = size, axis, None

    def qsize(self):
        return 0 if self.q is None or self.q.size == 0 else self.q.shape[self.axis]

    def put(self, a):
        self.q = a if self.qsize() == 0 else np.concatenate((self.q, a), self.axis)

    def get(self):
        out = _ax(self.q, 0, self.size, self.axis)
        self.q = _ax(self.q, self.size, None, self.axis)
        return out


# --------------------------------------------------------------------------
# FIR: overlap-add convolution   (core/numerical.py:158-298)
# --------------------------------------------------------------------------
def _optimal_nfft(ntaps):
    """core/numerical.py:19-38."""
    return int(8 * 2 ** np.ceil(np.log2(ntaps)))


def oaconvolve(x, window, chunksize, axis=-1, mode="same", nfft_factor=32):
    """Blocks yielded by ``nm.oaconvolve(producer(x, chunksize, axis), window,
    axis, mode, nfft_factor)``  (core/numerical.py:158-298).

    Per block: zero-pad K-1, rfft(nfft), multiply by H=rfft(window, nfft),
    irfft, split at ``step``, add the previous block's K-1 tail (:229-269).
    The first block is cut on the left and the last on the right by the numpy
    convolve mode (:119-155, :272-273, :298).
    """
    x = np.asarray(x, dtype=float)
    window = np.asarray(window, dtype=float)
    K, N = len(window), x.shape[axis]

    nfft = _optimal_nfft(K) * nfft_factor                       # :202
    if nfft - K + 1 > N:                                        # :210-211
        nfft = min(_optimal_nfft(K), N)
    H = np.fft.rfft(window, nfft)                               # :214
    step = nfft - K + 1                                         # :217

    hshape = [1] * x.ndim
    hshape[axis] = len(H)
    Hb = H.reshape(hshape)

    def cconv(seg):                                             # :229-241
        pads = [(0, 0)] * seg.ndim
        pads[axis] = (0, K - 1)
        padded = np.pad(seg, pads)
        spec = np.fft.rfft(padded, nfft, axis=axis) * Hb
        return np.fft.irfft(spec, axis=axis).real

    left_cut = {"full": 0, "same": (K - 1) // 2, "valid": K - 1}[mode]       # :143-150
    right_cut = {"full": 0, "same": int(np.ceil((K - 1) / 2)), "valid": K - 1}[mode]

    oshape = list(x.shape)
    oshape[axis] = K - 1
    overlap = np.zeros(oshape)                                  # :221-223
    fifo = _Fifo(step, axis)
    out, nseg = [], 0
    for chunk in split_chunks(x, chunksize, axis):              # :254
        fifo.put(chunk)
        while fifo.qsize() > step:                              # :258 (strict >)
            z = cconv(fifo.get())
            y = np.array(_ax(z, 0, step, axis))
            new_overlap = _ax(z, step, None, axis)
            sl = [slice(None)] * y.ndim
            sl[axis] = slice(0, K - 1)
            y[tuple(sl)] += overlap                             # :243-251
            overlap = new_overlap
            if nseg == 0:
                y = _ax(y, left_cut, None, axis)                # :272-273
            nseg += 1
            out.append(y)
    if fifo.qsize() > 0:                                        # :285-298
        tail = fifo.q
        z = cconv(tail)
        y = np.array(_ax(z, 0, tail.shape[axis] + K - 1, axis))
        sl = [slice(None)] * y.ndim
        sl[axis] = slice(0, K - 1)
        y[tuple(sl)] += overlap
        out.append(_ax(y, 0, y.shape[axis] - right_cut, axis))
    return out


# --------------------------------------------------------------------------
# IIR: second order sections  (core/numerical.py:301-411)
# --------------------------------------------------------------------------
def sosfilt(x, sos, chunksize, axis=-1, zi=None):
    """Forward SOS cascade with the delay registers carried from chunk to
    chunk (core/numerical.py:301-335).  Returns (chunks, final_state)."""
    x = np.asarray(x, dtype=float)
    sos = np.asarray(sos, dtype=float)
    zshape = list(x.shape)
    zshape[axis] = 2
    z = np.zeros((sos.shape[0], *zshape)) if zi is None else np.array(zi, dtype=float)
    out = []
    for chunk in split_chunks(x, chunksize, axis):
        y, z = sps.sosfilt(sos, chunk, axis=axis, zi=z)         # :334
        out.append(y)
    return out, z


def _zi_sos(sos, ndim, axis):
    """Steady-state state per unit input, broadcastable (numerical.py:378-382)."""
    zi = sps.sosfilt_zi(sos)
    s = [1] * ndim
    s[axis] = 2
    return zi.reshape((sos.shape[0], *s))


def sosfiltfilt(x, sos, chunksize, axis=-1):
    """Chunk-dependent forward-backward SOS filter (core/numerical.py:338-411).

    1. ONE forward pass over the whole signal, state carried across chunks,
       initial state zi*x[0]                                  (:374-386)
    2. chunk i < last: run the filter over flip(F[i+1]) from zi*F[i+1][-1],
       keep only the final state zf; output flip(sosfilt(flip(F[i]), zi=zf))
                                                               (:394-403)
    3. last chunk: backward pass from zi*F[last][-1]           (:408-411)
    """
    x = np.asarray(x, dtype=float)
    sos = np.asarray(sos, dtype=float)
    zi = _zi_sos(sos, x.ndim, axis)
    x0 = _ax(x, 0, 1, axis)
    fwd, _ = sosfilt(x, sos, chunksize, axis, zi=zi * x0)
    n = len(fwd)
    out = []
    for i, a in enumerate(fwd):
        af = np.flip(a, axis=axis)
        if i < n - 1:
            bf = np.flip(fwd[i + 1], axis=axis)
            _, zf = sps.sosfilt(sos, bf, axis=axis, zi=zi * _ax(bf, 0, 1, axis))
        else:
            zf = zi * _ax(af, 0, 1, axis)
        r, _ = sps.sosfilt(sos, af, axis=axis, zi=zf)
        out.append(np.flip(r, axis=axis))
    return out


# --------------------------------------------------------------------------
# IIR: transfer function (b, a)  (core/numerical.py:414-520)
# --------------------------------------------------------------------------
def lfilter(x, coeffs, chunksize, axis=-1, zi=None):
    """Forward (b, a) filter with carried state (core/numerical.py:414-446)."""
    x = np.asarray(x, dtype=float)
    b, a = coeffs
    zshape = list(x.shape)
    zshape[axis] = int(max(len(b), len(a)) - 1)                 # :439
    z = np.zeros(zshape) if zi is None else np.array(zi, dtype=float)
    out = []
    for chunk in split_chunks(x, chunksize, axis):
        y, z = sps.lfilter(b, a, chunk, axis=axis, zi=z)        # :445
        out.append(y)
    return out, z


def filtfilt(x, coeffs, chunksize, axis=-1):
    """Chunk-dependent forward-backward (b, a) filter (numerical.py:449-520);
    same three steps as :func:`sosfiltfilt` with ``lfilter_zi`` (:487-491)."""
    x = np.asarray(x, dtype=float)
    b, a = coeffs
    zi = sps.lfilter_zi(b, a)
    s = [1] * x.ndim
    s[axis] = zi.size
    zi = zi.reshape(s)
    fwd, _ = lfilter(x, coeffs, chunksize, axis, zi=zi * _ax(x, 0, 1, axis))
    n = len(fwd)
    out = []
    for i, f in enumerate(fwd):
        ff = np.flip(f, axis=axis)
        if i < n - 1:
            gf = np.flip(fwd[i + 1], axis=axis)
            _, zf = sps.lfilter(b, a, gf, axis=axis, zi=zi * _ax(gf, 0, 1, axis))
        else:
            zf = zi * _ax(ff, 0, 1, axis)
        r, _ = sps.lfilter(b, a, ff, axis=axis, zi=zf)
        out.append(np.flip(r, axis=axis))
    return out


# --------------------------------------------------------------------------
# polyphase resampling  (core/numerical.py:523-632)
# --------------------------------------------------------------------------
def _kaiser_lowpass(fpass, fstop, fs, gpass, gstop):
    """Kaiser-window FIR design used by the resampler: filtering/fir.py:122-137
    (numtaps, beta) + filtering/bases.py:347-361 (firwin), lowpass only."""
    pass_att = -20 * np.log10(1 - 10 ** (-gpass / 20))          # bases.py:305-309
    ripple = max(pass_att, gstop)
    width = abs(fstop - fpass)
    ntaps, _ = sps.kaiserord(ripple, width / (fs / 2))
    ntaps = ntaps + 1 if ntaps % 2 == 0 else ntaps
    cutoff = min(fpass, fstop) + width / 2                      # bases.py:311-316
    return sps.firwin(ntaps, cutoff=cutoff, width=None,
                      window=("kaiser", sps.kaiser_beta(ripple)),
                      pass_zero="lowpass", scale=True, fs=fs)


def resample_filter(L, M, fs, **kwargs):
    """Anti-alias / interpolation taps (core/numerical.py:579-583)."""
    cutoff = fs / (2 * max(L, M))
    fstop = kwargs.pop("fstop", cutoff + cutoff / 10)
    fpass = kwargs.pop("fpass", cutoff - cutoff / 10)
    gpass, gstop = kwargs.pop("gpass", 0.1), kwargs.pop("gstop", 40)
    return _kaiser_lowpass(fpass, fstop, fs, gpass, gstop)


def polyphase_resample(x, L, M, fs, chunksize, axis=-1, **kwargs):
    """Arrays yielded by ``nm.polyphase_resample`` (core/numerical.py:523-632).

    Chunk size is capped at N//3 and rounded up to a multiple of M
    (:574-587); each chunk is resampled with ``overhang`` samples of halo
    on both sides (zeros at the recording edges) and the halo's outputs are
    sliced off (:597-632); the last two chunks are merged (:625-628).
    """
    x = np.asarray(x, dtype=float)
    N = x.shape[axis]
    if M >= N:                                                  # :569-571
        raise ValueError("Decimation factor must M={} be < pro.shape[{}] = {}"
                         .format(M, axis, N))
    csize = int(chunksize)
    if csize > N // 3:                                          # :574-576
        csize = N // 3
    h = resample_filter(L, M, fs, **kwargs)                     # :579-583
    if csize % M > 0:                                           # :586-587
        csize = int(np.ceil(csize / M) * M)
    overhang = int(np.ceil((len(h) - 1) / M) * M)               # :597
    a = int(overhang * L / M)                                   # :613

    chunks = split_chunks(x, csize, axis)
    cnt = N // csize + bool(N % csize) - 1                      # :617
    zshape = list(x.shape)
    zshape[axis] = overhang
    zeros = np.zeros(zshape)

    out = []
    # first chunk: zeros | chunk0 | head of chunk1              (:600-614)
    right = _ax(chunks[1], 0, overhang, axis)
    padded = np.concatenate((zeros, chunks[0], right), axis=axis)
    r = sps.resample_poly(padded, up=L, down=M, axis=axis, window=h)
    out.append(_ax(r, a, -a, axis))
    # remaining: zip(iprior, icurrent, inext) runs while inext has data, i.e.
    # n = 1 .. len(chunks)-2                                    (:618-632)
    for n in range(1, len(chunks) - 1):
        left = _ax(chunks[n - 1], -overhang, None, axis)
        curr, nxt = chunks[n], chunks[n + 1]
        if n < cnt - 1:
            right = _ax(nxt, 0, overhang, axis)
        else:
            curr = np.concatenate((curr, nxt), axis=axis)
            right = np.zeros(left.shape)
        padded = np.concatenate((left, curr, right), axis=axis)
        r = sps.resample_poly(padded, L, M, axis=axis, window=h)
        out.append(_ax(r, a, -a, axis))
    return out


# --------------------------------------------------------------------------
# spectra  (core/numerical.py:635-1087, spectra/estimators.py:141-156,269-284)
# --------------------------------------------------------------------------
def modified_dft(arr, fs, nfft, window="hann", axis=-1, detrend="constant",
                 scaling="density"):
    """Windowed DFT of one segment (core/numerical.py:635-718)."""
    arr = np.asarray(arr, dtype=float)
    if nfft < arr.shape[axis]:
        arr = _ax(arr, 0, nfft, -1)                             # :688 (axis -1 quirk)
    arr = sps.detrend(arr, axis=axis, type=detrend)             # :691
    w = sps.get_window(window, arr.shape[axis])                 # :694
    s = [1] * arr.ndim
    s[axis] = len(w)
    arr = arr * w.reshape(s)
    X = np.fft.rfft(arr, nfft, axis=axis)                       # :699
    freqs = np.fft.rfftfreq(nfft, d=1 / fs)
    if scaling == "spectrum":                                   # :703-712
        norm = 1 / np.sum(w) ** 2
    elif scaling == "density":
        norm = 1 / (fs * np.sum(w ** 2))
    else:
        raise ValueError("Unknown scaling: {}".format(scaling))
    X = X * np.sqrt(norm)                                       # :716
    return freqs, X


def periodogram(arr, fs, nfft=None, window="hann", axis=-1, detrend="constant",
                scaling="density"):
    """|modified_dft|^2 with one-sided doubling (core/numerical.py:721-796)."""
    arr = np.asarray(arr, dtype=float)
    nfft = arr.shape[axis] if not nfft else int(nfft)
    freqs, X = modified_dft(arr, fs, nfft, window, axis, detrend, scaling)
    P = np.real(X) ** 2 + np.imag(X) ** 2                       # :782
    sl = [slice(None)] * P.ndim
    sl[axis] = slice(1, None) if nfft % 2 else slice(1, -1)     # :785-792
    P[tuple(sl)] *= 2
    return freqs, P


def segments(x, nfft, overlap, chunksize, axis=-1):
    """The nfft-long windows ``_spectra_estimatives`` hands to its ``func``
    (core/numerical.py:799-849): FIFO of the incoming chunks, window k starts
    at k*stride, stride = nfft - int(nfft*overlap); trailing partial dropped."""
    x = np.asarray(x, dtype=float)
    stride = nfft - int(nfft * overlap)                         # :817-818
    fifo = _Fifo(stride, axis)
    segs = []
    for chunk in split_chunks(x, chunksize, axis):
        while fifo.qsize() >= nfft:                             # :825-833
            segs.append(_ax(fifo.q, 0, nfft, axis))
            fifo.get()
        fifo.put(chunk)                                         # :835-839
    while fifo.qsize() >= nfft:                                 # :844-849
        segs.append(_ax(fifo.q, 0, nfft, axis))
        fifo.get()
    return segs


def welch_psd(x, fs, axis=-1, resolution=0.5, window="hann", overlap=0.5,
              detrend="constant", scaling="density"):
    """``spectra.estimators.psd`` (spectra/estimators.py:141-156): chunksize is
    forced to int(fs), nfft = int(fs/resolution), periodogram per window,
    incremental mean r += (p - r)/cnt.  Returns (cnt, freqs, estimate)."""
    x = np.asarray(x, dtype=float)
    nfft = int(fs / resolution)                                 # :144
    freqs = np.fft.rfftfreq(nfft, 1 / fs)
    result, cnt = 0, 0
    for cnt, seg in enumerate(segments(x, nfft, overlap, int(fs), axis), 1):
        _, p = periodogram(seg, fs, nfft, window, axis, detrend, scaling)
        result = result + 1 / cnt * (p - result)                # :150-152
    return cnt, freqs, result


def stft(x, fs, axis=-1, resolution=0.5, window="hann", overlap=0.5,
         detrend="constant", scaling="density", boundary=True, padded=True):
    """``spectra.estimators.stft`` with asarray=True (estimators.py:269-284 ->
    core/numerical.py:950-1087).  Returns (freqs, time, X) with the segments
    stacked on a new last axis."""
    x = np.asarray(x, dtype=float)
    nfft = int(fs / resolution)
    noverlap = int(nfft * overlap)
    stride = nfft - noverlap
    N = x.shape[axis]
    data = x
    pads = [(0, 0)] * x.ndim
    if boundary:                                                # :1041-1044
        pads[axis] = (nfft // 2, nfft // 2)
        data = np.pad(data, pads)
    if padded:                                                  # :1046-1051
        amt = stride if N % stride else 0
        pads[axis] = (0, amt)
        data = np.pad(data, pads)
    Nd = data.shape[axis]
    freqs = np.fft.rfftfreq(nfft, 1 / fs)
    if boundary:                                                # :1076-1083
        time = 1 / fs * np.arange(0, Nd - nfft + 1, nfft - noverlap)
    else:
        time = 1 / fs * np.arange(nfft // 2, Nd + 1 - nfft // 2, nfft - noverlap)
    cols = []
    for seg in segments(data, nfft, overlap, int(fs), axis):
        _, X = modified_dft(seg, fs, nfft, window, axis, detrend, scaling)
        cols.append(X)
    return freqs, time, np.stack(cols, axis=-1)


# --------------------------------------------------------------------------
# producer tools: masked producers and reductions
# --------------------------------------------------------------------------
def masked(x, mask, chunksize, axis=-1):
    """Arrays yielded by ``producer(x, chunksize, axis, mask=mask)``
    (core/producer.py:427-444): chunk k of the data is paired with chunk k of
    the mask (``zip``: production stops when either runs out), chunks without a
    kept sample are skipped, the kept samples (``np.take`` of ``flatnonzero``)
    are re-chunked to ``chunksize`` by a FIFO."""
    x = np.asarray(x)
    mask = np.asarray(mask)
    fifo, out = _Fifo(int(chunksize), axis), []
    for arr, keep in zip(split_chunks(x, chunksize, axis), split_chunks(mask, chunksize, 0)):
        if not np.any(keep):                                    # :432-433
            continue
        fifo.put(np.take(arr, np.flatnonzero(keep), axis=axis))  # :436-437
        while fifo.qsize() >= fifo.size:                        # :439-441
            out.append(fifo.get())
    if fifo.qsize() > 0:                                        # :444-446
        out.append(fifo.get())
    return out


def pro_mean(chunks, pro_axis, axis=-1, ignore_nan=True, keepdims=False):
    """``protools.mean`` (core/protools.py:500-543) over the list of arrays a
    producer yields along ``pro_axis``."""
    averager = np.nanmean if ignore_nan else np.mean
    ndim = chunks[0].ndim
    ax = int(np.arange(ndim)[axis])
    if ax == int(np.arange(ndim)[pro_axis]):
        sums, cnts = 0, 0
        for arr in chunks:                                      # :531-534
            cnts += arr.shape[axis]
            sums += arr.shape[axis] * averager(arr, axis=axis, keepdims=keepdims)
        return sums / cnts                                      # :536
    avgs = [averager(a, axis=ax, keepdims=True) for a in chunks]   # :539
    result = np.concatenate(avgs, axis=pro_axis)
    return result if keepdims else np.squeeze(result, ax)


def pro_std(chunks, pro_axis, axis=-1, ignore_nan=True, keepdims=False):
    """``protools.std`` (core/protools.py:546-595): sqrt(E[x^2] - E[x]^2) with
    the chunk-weighted means of ``pro_mean``."""
    averager = np.nanmean if ignore_nan else np.mean
    dev = np.nanstd if ignore_nan else np.std
    ndim = chunks[0].ndim
    ax = int(np.arange(ndim)[axis])
    if ax == int(np.arange(ndim)[pro_axis]):
        expected_squared = pro_mean(chunks, pro_axis, ax, ignore_nan, keepdims) ** 2   # :580
        sum_squares, cnts = 0, 0
        for arr in chunks:                                      # :582-587
            cnts += arr.shape[axis]
            sum_squares += arr.shape[axis] * averager(arr ** 2, axis=axis, keepdims=keepdims)
        return np.sqrt(sum_squares / cnts - expected_squared)   # :590
    stds = [dev(a, axis=ax, keepdims=True) for a in chunks]     # :593
    result = np.concatenate(stds, axis=pro_axis)
    return result if keepdims else np.squeeze(result, ax)


def standardize(chunks, pro_axis, axis=-1, ignore_nan=True):
    """Arrays yielded by ``protools.standardize`` (core/protools.py:598-668)."""
    means = pro_mean(chunks, pro_axis, axis, ignore_nan, keepdims=True)    # :633
    stds = pro_std(chunks, pro_axis, axis, ignore_nan, keepdims=True)      # :634
    ndim = chunks[0].ndim
    if int(np.arange(ndim)[axis]) == int(np.arange(ndim)[pro_axis]):
        return [(arr - means) / stds for arr in chunks]         # :659-662
    out, pos = [], 0
    for arr in chunks:                                          # :664-668
        n = arr.shape[pro_axis]
        out.append((arr - _ax(means, pos, pos + n, pro_axis)) / _ax(stds, pos, pos + n, pro_axis))
        pos += n
    return out
