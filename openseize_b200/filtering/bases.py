"""FIR / IIR operator bases: design on the host with scipy, apply on the GPU.

The call signatures are the reference's, verbatim (filtering/bases.py:153-213
for IIR, :363-421 for FIR): ndarray in -> ndarray out, Producer in -> lazy
Producer out, built as ``producer(partial(nm.<genfunc>, pro, ...), ...)``.
"""

import abc
from functools import partial

import numpy as np
import scipy.signal as sps

from openseize_b200.core import numerical as nm
from openseize_b200.core.producer import producer


def _band_type(fpass, fstop, multiband=False):
    if len(fpass) < 2:
        return "lowpass" if fpass < fstop else "highpass"
    if len(fpass) == 2 or not multiband:
        return "bandstop" if fpass[0] < fstop[0] else "bandpass"
    return "multiband"


def _check_bands(fpass, fstop):
    if len(fpass) != len(fstop):
        raise ValueError("fpass and fstop must have the same shape, got {} and {}"
                         .format(fpass.shape, fstop.shape))


class IIR(abc.ABC):
    """Infinite impulse response filter designed by ``scipy.signal.iirfilter``
    (reference filtering/bases.py:19-213)."""

    def __init__(self, fpass, fstop, gpass, gstop, fs, fmt):
        self.fs = fs
        self.nyq = fs / 2
        self.fpass = np.atleast_1d(fpass)
        self.fstop = np.atleast_1d(fstop)
        _check_bands(self.fpass, self.fstop)
        self.gpass = gpass
        self.gstop = gstop
        self.fmt = "sos" if fmt == "zpk" else fmt
        self.coeffs = self._build()

    @property
    def ftype(self):
        return type(self).__name__.lower()

    @property
    def btype(self):
        return _band_type(self.fpass, self.fstop)

    @property
    @abc.abstractmethod
    def order(self):
        """(order, critical frequency) of the design."""

    def _build(self):
        n, wn = self.order
        return sps.iirfilter(n, wn, rp=self.gpass, rs=self.gstop, btype=self.btype,
                             ftype=self.ftype, output=self.fmt, fs=self.fs)

    def __call__(self, data, chunksize, axis=-1, dephase=True, zi=None, **kwargs):
        pro = producer(data, chunksize, axis, **kwargs)
        if self.fmt == "sos":
            if dephase:
                genfunc = partial(nm.sosfiltfilt, pro, self.coeffs, axis)
            else:
                genfunc = partial(nm.sosfilt, pro, self.coeffs, axis, zi)
        elif self.fmt == "ba":
            if dephase:
                genfunc = partial(nm.filtfilt, pro, self.coeffs, axis)
            else:
                genfunc = partial(nm.lfilter, pro, self.coeffs, axis, zi)
        else:
            raise ValueError("unknown coefficient format {!r}".format(self.fmt))
        result = producer(genfunc, chunksize, axis, shape=pro.shape)
        if isinstance(data, np.ndarray):
            result = result.to_array()
        return result


class FIR(abc.ABC):
    """Windowed-sinc finite impulse response filter designed by
    ``scipy.signal.firwin`` (reference filtering/bases.py:216-421)."""

    def __init__(self, fpass, fstop, gpass, gstop, fs, **kwargs):
        self.fpass = np.atleast_1d(fpass)
        self.fstop = np.atleast_1d(fstop)
        _check_bands(self.fpass, self.fstop)
        self.gpass = gpass
        self.gstop = gstop
        self.fs = fs
        self.nyq = fs / 2
        self.width = np.min(np.abs(self.fstop - self.fpass))
        self.coeffs = self._build(**kwargs)

    @property
    def ftype(self):
        return type(self).__name__.lower()

    @property
    def btype(self):
        if len(self.fpass) > 2:
            raise ValueError("{} supports only lowpass, highpass, bandpass & bandstop."
                             .format(type(self)))
        return _band_type(self.fpass, self.fstop)

    @property
    def pass_attenuation(self):
        """Pass band ripple expressed as an attenuation in dB."""
        return -20 * np.log10(1 - 10 ** (-self.gpass / 20))

    @property
    def cutoff(self):
        """-6 dB point of each transition band."""
        delta = abs(self.fstop - self.fpass) / 2
        return delta + np.min(np.stack((self.fpass, self.fstop)), axis=0)

    @property
    def window_params(self):
        return tuple()

    @property
    @abc.abstractmethod
    def numtaps(self):
        """Taps needed to meet the attenuation criteria in the transition width."""

    def _build(self, **kwargs):
        window = (self.ftype, *self.window_params)
        return sps.firwin(self.numtaps, cutoff=self.cutoff, width=None, window=window,
                          pass_zero=self.btype, scale=True, fs=self.fs, **kwargs)

    def __call__(self, data, chunksize, axis=-1, mode="same", **kwargs):
        pro = producer(data, chunksize, axis, **kwargs)
        genfunc = partial(nm.oaconvolve, pro, self.coeffs, axis, mode)
        shape = nm.convolved_shape(data.shape, self.coeffs.shape, mode, axis)
        result = producer(genfunc, chunksize, axis, shape=shape)
        if isinstance(data, np.ndarray):
            result = result.to_array()
        return result
