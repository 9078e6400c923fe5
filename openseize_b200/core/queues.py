"""FIFO re-chunking queue (reference: core/queues.py:9-70).

Same contract as the reference's ``FIFOArray`` -- ``put`` appends along
``axis``, ``get`` pops ``chunksize`` samples, ``queue`` is the concatenated
content -- but blocks are kept in a list and only concatenated when a ``get``
straddles them, so a stream of already chunk-sized blocks passes through
without a host copy."""

import numpy as np

from openseize_b200.core.arraytools import slice_along_axis


class FIFOArray:
    def __init__(self, chunksize, axis):
        self.chunksize = int(chunksize)
        self.axis = axis
        self._blocks = []
        self._size = 0

    # -- reference-compatible surface --------------------------------------
    @property
    def queue(self):
        if not self._blocks:
            return np.array([])
        if len(self._blocks) > 1:
            self._blocks = [np.concatenate(self._blocks, axis=self.axis)]
        return self._blocks[0]

    @queue.setter
    def queue(self, value):
        value = np.asarray(value)
        self._blocks = [value] if value.size else []
        self._size = value.shape[self.axis] if value.size else 0

    def qsize(self):
        return self._size

    def empty(self):
        return self._size == 0

    def full(self):
        return self._size >= self.chunksize

    def put(self, x):
        if x.size == 0:
            return
        self._blocks.append(x)
        self._size += x.shape[self.axis]

    def get(self):
        """Pop ``chunksize`` samples (fewer if the queue holds fewer)."""
        want = min(self.chunksize, self._size)
        taken, parts = 0, []
        while taken < want:
            blk = self._blocks[0]
            n = blk.shape[self.axis]
            if taken + n <= want:
                parts.append(blk)
                self._blocks.pop(0)
                taken += n
            else:
                k = want - taken
                parts.append(slice_along_axis(blk, 0, k, axis=self.axis))
                self._blocks[0] = slice_along_axis(blk, k, None, axis=self.axis)
                taken = want
        self._size -= want
        if not parts:
            return np.array([])
        return parts[0] if len(parts) == 1 else np.concatenate(parts, axis=self.axis)
