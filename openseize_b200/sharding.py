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
  bound).

Everything here is host-side planning plus one ``all_reduce``; the arithmetic
is the same GPU kernels (``numerical.welch_sum``).
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


def shard_channels(data, chunksize, axis=-1, group=None):
    """Producer over this rank's block of channels of an ndarray (a view -- no
    copy).  Running any operator chain on it and concatenating the ranks'
    results along the split axis equals the single-process result."""
    if isinstance(data, Producer):
        if not isinstance(data, ArrayProducer):
            raise TypeError("channel sharding slices in-memory data; shard generator or "
                            "reader producers where they are built")
        data = data.data
    rank, size = world(group)
    index, _ = channel_block(data.shape, axis, rank, size)
    return producer(data[index], chunksize, axis)


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
    if isinstance(data, ArrayProducer):
        data = data.data
    if not isinstance(data, np.ndarray):
        raise TypeError("time sharding slices in-memory data")
    axis = normalize_axis(axis, data.ndim)
    rank, size = world(group)
    nfft = int(fs / resolution)
    nsamples = data.shape[axis]
    start, stop = time_span(nsamples, nfft, overlap, rank, size)
    layout = dv.Layout(data.shape, axis)
    if stop > start:
        mine = slice_along_axis(data, start, stop, axis=axis)
        cnt, psd_sum = nm.welch_sum(producer(mine, int(fs), axis), fs, nfft, window, overlap,
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
