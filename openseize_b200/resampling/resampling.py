"""downsample / upsample / resample with the reference's signatures
(resampling/resampling.py:95-311) over the GPU polyphase kernel."""

from functools import partial

import numpy as np

from openseize_b200.core.numerical import polyphase_resample
from openseize_b200.core.producer import producer
from openseize_b200.filtering.fir import Kaiser


def resampled_shape(pro, L, M, axis):
    """Shape after resampling by L/M along ``axis``: ceil(N*L/M)."""
    shape = list(pro.shape)
    shape[axis] = int(np.ceil(pro.shape[axis] * L / M))
    return tuple(shape)


def _resample(data, L, M, fs, chunksize, axis, shape_lm, kwargs):
    pro = producer(data, chunksize, axis)
    genfunc = partial(polyphase_resample, pro, L, M, fs, Kaiser, axis, **kwargs)
    shape = resampled_shape(pro, L=shape_lm[0], M=shape_lm[1], axis=axis)
    result = producer(genfunc, chunksize, axis, shape=shape)
    return result.to_array() if isinstance(data, np.ndarray) else result


def downsample(data, M, fs, chunksize, axis=-1, **kwargs):
    """Decimate by the integer factor M with a Kaiser anti-alias filter."""
    if M == 1:
        return data
    return _resample(data, 1, M, fs, chunksize, axis, (1, M), kwargs)


def upsample(data, L, fs, chunksize, axis=-1, **kwargs):
    """Interpolate by the integer factor L with a Kaiser interpolation filter."""
    if L == 1:
        return data
    return _resample(data, L, 1, fs, chunksize, axis, (L, 1), kwargs)


def resample(data, L, M, fs, chunksize, axis=-1, **kwargs):
    """Resample by the rational factor L/M (reduced by their gcd)."""
    g = np.gcd(L, M)
    l, m = L // g, M // g
    if l == m == 1:
        return data
    return _resample(data, int(l), int(m), fs, chunksize, axis, (L, M), kwargs)
