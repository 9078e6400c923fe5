"""Producer reductions on the GPU, with the names and signatures of the
reference's ``openseize.core.protools`` (:500-668): ``mean``, ``std`` and
``standardize`` of a producer's values along an axis (SURVEY.md section 8f, N3
-- the quickstart's state-masked, standardized PSD workflow).

The producer's chunks are reduced on the device (``csrc/protools.cu``): host
producers are streamed through pinned memory, producers over GPU generating
functions (filters, resamplers, masked producers over them) are consumed in
HBM.  ``standardize`` returns a producer whose generating function has a device
twin, so a GPU operator downstream takes the standardized chunks without a
host round trip.  There is no CPU arithmetic here beyond the final
``sums / counts`` on ``rows`` numbers.

The structural helpers of the reference module (squeeze, add, multiply, pad,
expand_dims, slice_along_axis) are host-side generator plumbing and are out of
scope.
"""

import functools

import numpy as np

from openseize_b200.core import device as dv
from openseize_b200.core import numerical as nm
from openseize_b200.core.arraytools import normalize_axis
from openseize_b200.core.producer import producer


def _row_sums(pro, ignore_nan):
    """(rows, 3) host array: sum over chunks of n * chunk mean, of n * chunk mean
    of squares, and of n -- the reference's chunk-weighted sums
    (protools.py:531-536, 583-590), one pass for both moments."""
    dv.require_cuda()
    layout = dv.Layout(pro.shape, pro.axis)
    acc = dv.RowMoments(layout.rows, ignore_nan)
    for block in nm.device_chunks(pro, pro.axis, regrid=True):
        acc.add(block)
    return acc.result(), layout


def _shape_result(values, layout, axis, keepdims):
    res = np.asarray(values, dtype=np.float64).reshape(layout.host_shape(1))
    if not keepdims:
        res = np.squeeze(res, axis)
    return res[()] if res.ndim == 0 else res


def _col_stat(pro, ax, ignore_nan, keepdims, want):
    """Statistic across a non-production axis: every chunk on its own
    (protools.py:538-543, 592-595)."""
    if pro.ndim != 2:
        raise NotImplementedError(
            "protools.{} across a non-production axis is built for 2-D producers".format(want))
    shape = list(pro.shape)
    shape[ax] = 1
    out_layout = dv.Layout(shape, pro.axis)
    blocks = (dv.col_moments(block, ignore_nan, want)
              for block in nm.device_chunks(pro, pro.axis, regrid=True))
    result = np.concatenate(list(nm._to_host(blocks, out_layout)), axis=pro.axis)
    return result if keepdims else np.squeeze(result, ax)


def mean(pro, axis=-1, ignore_nan=True, keepdims=False):
    """Mean of a producer's values along ``axis``; across all produced arrays
    when ``axis`` is the production axis (reference protools.py:500-543)."""
    ax = normalize_axis(axis, pro.ndim)
    if ax != pro.axis:
        return _col_stat(pro, ax, ignore_nan, keepdims, "mean")
    acc, layout = _row_sums(pro, ignore_nan)
    return _shape_result(acc[:, 0] / acc[:, 2], layout, ax, keepdims)


def std(pro, axis=-1, ignore_nan=True, keepdims=False):
    """Standard deviation sqrt(E[x^2] - E[x]^2) along ``axis`` (reference
    protools.py:546-595)."""
    ax = normalize_axis(axis, pro.ndim)
    if ax != pro.axis:
        return _col_stat(pro, ax, ignore_nan, keepdims, "std")
    acc, layout = _row_sums(pro, ignore_nan)
    expected_squared = (acc[:, 0] / acc[:, 2]) ** 2
    return _shape_result(np.sqrt(acc[:, 1] / acc[:, 2] - expected_squared), layout, ax, keepdims)


def _standardize_device(pro, means, stds, axis, ignore_nan=True, _out=None, _free=False):
    dv.require_cuda()
    ax = normalize_axis(axis, pro.ndim)
    rows = dv.Layout(pro.shape, pro.axis).rows
    chunks = nm.device_chunks(pro, pro.axis, regrid=not _free)
    if ax != pro.axis:
        for block in chunks:
            yield dv.col_moments(block, ignore_nan, "standardize")
        return
    mu = dv.from_host(np.asarray(means, dtype=np.float64).reshape(-1))
    sd = dv.from_host(np.asarray(stds, dtype=np.float64).reshape(-1))
    for block in chunks:
        out = _out(rows, block.shape[1]) if _out is not None else None
        yield dv.row_standardize(block, mu, sd, out=out)


def _standardize_layout(pro, means, stds, axis, ignore_nan=True):
    return dv.Layout(pro.shape, pro.axis)


@nm._gpu_genfunc(_standardize_device, _standardize_layout)
def _standardize_gen(pro, means, stds, axis, ignore_nan=True):
    """(arr - means) / stds for every produced array (reference protools.py:
    640-668); across a non-production axis each chunk is standardized with its
    own per-sample statistics."""


def standardize(pro, axis=-1, ignore_nan=True):
    """Producer of the values of ``pro`` standardized along ``axis`` (reference
    protools.py:598-637).  As in the reference the mean and standard deviation
    along the production axis are computed when this function is called (one
    pass over ``pro`` here, three there)."""
    ax = normalize_axis(axis, pro.ndim)
    means = stds = None
    if ax == pro.axis:
        acc, layout = _row_sums(pro, ignore_nan)
        m = acc[:, 0] / acc[:, 2]
        means = m.reshape(layout.host_shape(1))
        stds = np.sqrt(acc[:, 1] / acc[:, 2] - m ** 2).reshape(layout.host_shape(1))
    elif pro.ndim != 2:
        raise NotImplementedError(
            "protools.standardize across a non-production axis is built for 2-D producers")
    func = functools.partial(_standardize_gen, pro, means, stds, axis, ignore_nan)
    return producer(func, pro.chunksize, pro.axis, shape=pro.shape)
