"""Producers: re-iterable objects that yield ndarrays of ``chunksize`` samples
along one axis.  Same contract as the reference's ``core/producer.py``
(:54-444): the ``producer`` factory dispatches on the data type, a Producer
passed back in is MUTATED (chunksize / axis) and returned (:114-117),
generating functions need an explicit shape (:318-320), and everything stays
picklable (tests/test_concurrency.py) because no CUDA state is created before
iteration starts.

Producers built from this package's GPU generating functions
(``openseize_b200.core.numerical``) can additionally be consumed on the device
by a downstream GPU operator without a host round trip -- see
``numerical.device_chunks``.
"""

import functools
import inspect
from collections import abc

import numpy as np

from openseize_b200.core import resources
from openseize_b200.core.arraytools import normalize_axis, slice_along_axis
from openseize_b200.core.queues import FIFOArray


def _is_reader(obj):
    """Duck-typed file reader (reference: file_io/bases.py Reader ABC)."""
    return all(hasattr(obj, name) for name in ("read", "shape", "open", "close")) \
        and not isinstance(obj, np.ndarray)


def producer(data, chunksize, axis, shape=None, mask=None, **kwargs):
    """Build a producer of ``chunksize``-long ndarrays along ``axis`` from an
    ndarray, a sequence of ndarrays, a reader, a generating function (needs
    ``shape``) or another producer.  With ``mask`` the produced samples are
    filtered by a 1-D boolean array along ``axis``."""
    if isinstance(data, Producer):
        data.chunksize = int(chunksize)
        data.axis = normalize_axis(axis, len(data.shape))
        result = data
    elif _is_reader(data):
        result = ReaderProducer(data, chunksize, axis=1, **kwargs)
    elif inspect.isgeneratorfunction(data) or (
            isinstance(data, functools.partial) and inspect.isgeneratorfunction(data.func)):
        if shape is None:
            raise ValueError("A Producer from a generating function requires a shape.")
        result = GenProducer(data, chunksize, normalize_axis(axis, len(shape)), shape, **kwargs)
    elif isinstance(data, np.ndarray):
        result = ArrayProducer(data, chunksize, normalize_axis(axis, data.ndim), **kwargs)
    elif isinstance(data, abc.Sequence):
        joined = np.concatenate(data, axis)
        result = ArrayProducer(joined, chunksize, normalize_axis(axis, joined.ndim), **kwargs)
    else:
        raise TypeError("unproducible type: {}".format(type(data)))
    if mask is None:
        return result
    return MaskedProducer(result, mask, chunksize, result.axis, **kwargs)


class Producer(abc.Iterable):
    """Base of all producers: ``data``, ``chunksize``, ``axis``, ``shape``."""

    def __init__(self, data, chunksize, axis, **kwargs):
        self.data = data
        self._chunksize = int(chunksize)
        self.axis = axis
        self.kwargs = kwargs

    @property
    def chunksize(self):
        return self._chunksize

    @chunksize.setter
    def chunksize(self, value):
        self._chunksize = int(value)

    @property
    def shape(self):
        raise NotImplementedError

    @property
    def ndim(self):
        return len(self.shape)

    def to_array(self, dtype=float, limit=None):
        """Concatenate every produced array along ``axis`` if it fits in RAM."""
        if resources.assignable(self.shape, dtype, limit=limit):
            return np.concatenate(list(self), axis=self.axis)
        return None

    def __repr__(self):
        return "{}(shape={}, chunksize={}, axis={})".format(
            type(self).__name__, self.shape, self.chunksize, self.axis)


class ArrayProducer(Producer):
    """Views of an in-memory ndarray."""

    @property
    def shape(self):
        return self.data.shape

    def __iter__(self):
        yield from self.blocks(self.chunksize)

    def blocks(self, step):
        """Views of ``step`` samples (a multiple of chunksize for consumers that
        do not depend on the chunk grid and want fewer, larger uploads)."""
        n = self.data.shape[self.axis]
        for start in range(0, n, step):
            yield slice_along_axis(self.data, start, min(start + step, n), axis=self.axis)


class ReaderProducer(Producer):
    """Chunks read from a file reader with ``read(start, stop)``; the reader is
    closed until iteration so the producer pickles."""

    def __init__(self, data, chunksize, axis, **kwargs):
        super().__init__(data, chunksize, axis, **kwargs)
        first = self.kwargs.pop("start", 0)
        last = self.kwargs.pop("stop", self.data.shape[axis])
        self.start, self.stop, _ = slice(first, last).indices(data.shape[axis])
        self.data.close()

    @property
    def shape(self):
        s = list(self.data.shape)
        s[self.axis] = self.stop - self.start
        return tuple(s)

    def __iter__(self):
        yield from self.blocks(self.chunksize)

    def blocks(self, step):
        self.data.open()
        for a in range(self.start, self.stop, step):
            yield self.data.read(a, min(a + step, self.stop), **self.kwargs)

    def iter_raw(self, step=None):
        """Chunks as the reader's raw records (``file_io.edf.RawChunk``: int16
        samples + calibration) for consumers that decode on the GPU; None when
        the reader cannot provide them (no ``read_raw``, mixed sample rates, a
        sample axis other than the last, extra ``read`` arguments)."""
        reader = self.data
        if (not hasattr(reader, "read_raw") or self.kwargs or len(reader.shape) != 2
                or self.axis not in (1, -1) or not getattr(reader, "uniform_rate", False)):
            return None

        step = int(step or self.chunksize)

        def chunks():
            reader.open()
            for a in range(self.start, self.stop, step):
                yield reader.read_raw(a, min(a + step, self.stop))

        return chunks()


class GenProducer(Producer):
    """Re-chunks whatever a generating function yields to ``chunksize``."""

    def __init__(self, data, chunksize, axis, shape, **kwargs):
        if shape is None:
            raise ValueError("A Producer from a generating function requires a shape.")
        super().__init__(data, chunksize, axis, **kwargs)
        self._shape = tuple(int(s) for s in shape)

    @property
    def shape(self):
        return self._shape

    def __iter__(self):
        fifo = FIFOArray(self.chunksize, self.axis)
        for block in self.data(**self.kwargs):
            fifo.put(block)
            while fifo.full():
                yield fifo.get()
        if not fifo.empty():
            yield fifo.get()


class MaskedProducer(Producer):
    """Samples of a producer selected by a boolean mask along ``axis``;
    production stops when the producer or the mask runs out."""

    def __init__(self, pro, mask, chunksize, axis, **kwargs):
        super().__init__(pro, chunksize, axis, **kwargs)
        self.mask = producer(np.asarray(mask), chunksize, axis=0)

    @property
    def shape(self):
        result = list(self.data.shape)
        kept = int(np.count_nonzero(self.mask.data[: self.data.shape[self.axis]]))
        result[self.axis] = kept
        return tuple(result)

    @property
    def chunksize(self):
        return self.data.chunksize

    @chunksize.setter
    def chunksize(self, value):
        self.data.chunksize = int(value)
        self.mask.chunksize = int(value)

    def __iter__(self):
        fifo = FIFOArray(self.chunksize, self.axis)
        for arr, keep in zip(self.data, self.mask):
            if not np.any(keep):
                continue
            fifo.put(np.take(arr, np.flatnonzero(keep), axis=self.axis))
            while fifo.full():
                yield fifo.get()
        if not fifo.empty():
            yield fifo.get()


class DeviceProducer(Producer):
    """Chunks that already live on the GPU (device-resident recordings, and
    bench.py's HBM-resident timing mode).  ``data`` is a callable returning an
    iterator of float64 CUDA tensors of shape (rows, n) -- time-contiguous
    rows, the package's device layout; ``shape`` is the (rows, total) shape.
    GPU operators consume it without any host traffic; iterating it on the
    host downloads each chunk."""

    def __init__(self, data, chunksize, shape, **kwargs):
        super().__init__(data, chunksize, axis=len(shape) - 1, **kwargs)
        self._shape = tuple(int(s) for s in shape)
        if len(self._shape) != 2:
            raise ValueError("DeviceProducer holds 2-D (rows, samples) data")

    @property
    def shape(self):
        return self._shape

    def device_iter(self):
        return iter(self.data(**self.kwargs))

    def __iter__(self):
        for block in self.device_iter():
            yield block.cpu().numpy()
