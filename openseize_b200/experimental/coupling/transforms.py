"""Transforms that turn real data into complex signals for amplitude and phase
extraction (reference experimental/coupling/transforms.py).  ``Analytic`` -- the
analytic signal x + i H(x) with the Type III FIR Hilbert transformer -- runs on the
GPU: the Hilbert FIR is ``nm.oaconvolve`` (mode 'same') and the complex signal is
assembled on the device, so one pass over the source yields complex128 chunks (the
reference iterates the source twice: ``protools.add(real, protools.multiply(imag, 1j))``,
transforms.py:185-192).  Amplitudes and phases are taken on the host from the complex
chunks exactly as the reference does (:57-66, :84-94).
"""

import abc
import functools

import numpy as np

from openseize_b200.core import device as dv
from openseize_b200.core import numerical as nm
from openseize_b200.core.producer import Producer, producer
from openseize_b200.filtering.special import Hilbert


def _analytic_device(pro, taps, axis, _out=None, _free=False):
    """Device rows (rows, n, 2) of x + 1j * convolve(x, taps, 'same'), one block per
    input chunk.  The FIR loop is nm.oaconvolve's; the samples of x that pair with each
    'same'-mode output block are kept in a second ring."""
    dv.require_cuda()
    taps = np.asarray(taps, dtype=np.float64)
    ntaps, nsamp = len(taps), pro.shape[axis]
    if nsamp < ntaps:
        raise ValueError("Analytic: data length {} along axis is shorter than the {} taps "
                         "of the Hilbert filter".format(nsamp, ntaps))
    left, right = nm._mode_cuts(ntaps, "same")
    plan = dv.FirPlan.cached(taps)
    rows = nm._layout_of(pro, axis).rows
    ring = nm._TimeRing(rows)                      # FIR halo, as in nm._oaconvolve_device
    ring.push_zeros(ntaps - 1)
    real = nm._TimeRing(rows)                      # x not yet paired; starts at sample `done`
    pos, seen, done = 0, 0, 0
    for chunk in nm.device_chunks(pro, axis, regrid=False, alloc=ring):
        n = chunk.shape[1]
        ring.push(chunk)
        real.push(chunk)
        seen += n
        final = seen >= nsamp
        if final:
            ring.push_zeros(ntaps - 1)
        n_out = n + (ntaps - 1 if final else 0)
        y = plan.run(ring.window(), n_out)
        ring.drop(n_out)
        # full-convolution indices [pos, pos + n_out) -> 'same' outputs [lo, hi) of this block
        lo = max(left - pos, 0)
        hi = min(pos + n_out, nsamp + left) - pos
        pos += n_out
        if hi > lo:
            m = hi - lo                            # 'same' output samples done .. done + m
            z = dv.zip_complex(real.window()[:, :m], y[:, lo:hi])
            real.drop(m)
            done += m
            yield z
        if final:
            break


def _analytic_layout(pro, taps, axis):
    return nm._layout_of(pro, axis)


def analytic_signal(pro, taps, axis):
    """Generating function of the complex analytic signal's chunks (host complex128)."""
    layout = _analytic_layout(pro, taps, axis)
    blocks = _analytic_device(pro, taps, axis)
    cs = int(getattr(pro, "chunksize", 0) or 0)
    if cs > 0:
        blocks = _regrid_complex(blocks, cs)
    yield from nm._to_host(blocks, layout, complex_=True)


def _regrid_complex(blocks, cs):
    """Re-block (rows, n, 2) device tensors to ``cs`` samples along time."""
    t = dv.torch()
    held, size = [], 0
    for b in blocks:
        held.append(b)
        size += b.shape[1]
        while size >= cs:
            buf = held[0] if len(held) == 1 else t.cat(held, dim=1)
            yield buf[:, :cs].contiguous()
            rest = buf[:, cs:]
            held, size = ([rest] if rest.shape[1] else []), rest.shape[1]
    if size:
        yield (held[0] if len(held) == 1 else t.cat(held, dim=1)).contiguous()


analytic_signal.device = _analytic_device


class Transform(abc.ABC):
    """Abstract base of the transforms (reference transforms.py:18-104): holds the raw
    ``data`` producer and the complex ``signal`` producer its ``estimate`` returns;
    ``amplitudes`` and ``phases`` are producers derived from the signal."""

    def __init__(self, data, fs, chunksize=int(10e6), axis=-1, **kwargs):
        self.fs = fs
        self.chunksize = chunksize
        self.axis = axis
        self.data = producer(data, chunksize, axis)
        self.signal = self.estimate(self.data, **kwargs)

    @abc.abstractmethod
    def estimate(self, data, **kwargs):
        """Returns a complex producer."""

    def _envelope(self):
        for arr in self.signal:
            yield np.abs(arr)

    @property
    def amplitudes(self):
        return producer(self._envelope, self.chunksize, self.axis, shape=self.signal.shape)

    def _phase(self):
        for arr in self.signal:
            phi = np.angle(arr)
            phi[phi < 0] += 2 * np.pi
            yield phi

    @property
    def phases(self):
        return producer(self._phase, self.chunksize, self.axis, shape=self.signal.shape)


class Analytic(Transform):
    """The Hilbert analytic transform x + i H(x) with the Type III FIR Hilbert
    transformer (reference transforms.py:107-192)."""

    def estimate(self, data, *, width, gpass=0.01, gstop=60, **kwargs):
        hilbert = Hilbert(width, fs=self.fs, gpass=gpass, gstop=gstop)
        real = producer(data, self.chunksize, self.axis)
        assert isinstance(real, Producer)
        genfunc = functools.partial(analytic_signal, real, hilbert.coeffs, self.axis)
        return producer(genfunc, self.chunksize, self.axis, shape=real.shape)
