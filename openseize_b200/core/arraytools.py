"""Axis-generic ndarray helpers used by the host side of the hot path
(reference: core/arraytools.py:6-82,118-134 -- normalize / pad / slice /
split / multiply along one axis).  On the device these are fused into the
kernels; the host versions serve producers and tests."""

import numpy as np


def normalize_axis(axis, ndim):
    """Non-negative index of ``axis`` in an ``ndim``-dimensional array."""
    axis = int(axis)
    if not -ndim <= axis < ndim:
        raise IndexError("axis {} is out of bounds for {} dimensions".format(axis, ndim))
    return axis % ndim


def _slicer(ndim, axis, sl):
    idx = [slice(None)] * ndim
    idx[axis] = sl
    return tuple(idx)


def slice_along_axis(arr, start=None, stop=None, step=None, axis=-1):
    return arr[_slicer(arr.ndim, axis, slice(start, stop, step))]


def split_along_axis(arr, index, axis=-1):
    return (slice_along_axis(arr, 0, index, axis=axis),
            slice_along_axis(arr, index, None, axis=axis))


def pad_along_axis(arr, pad, axis=-1, **kwargs):
    pad = [pad, pad] if isinstance(pad, int) else list(pad)
    widths = [(0, 0)] * arr.ndim
    widths[axis] = tuple(pad)
    return np.pad(arr, widths, **kwargs)


def multiply_along_axis(x, y, axis=-1):
    shape = [1] * x.ndim
    shape[axis] = len(y)
    return x * np.reshape(y, shape)
