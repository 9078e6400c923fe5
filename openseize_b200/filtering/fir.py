"""Concrete FIR designs (reference filtering/fir.py:52-664): Kaiser and the
fixed windows via ``firwin``, Parks-McClellan via ``remez``.  Design is host
side scipy; application is ``FIR.__call__`` -> GPU ``nm.oaconvolve``."""

import numpy as np
import scipy.signal as sps

from openseize_b200.filtering.bases import FIR


def _odd(ntaps):
    """Odd tap count so the group delay is a whole number of samples."""
    return ntaps + 1 if ntaps % 2 == 0 else ntaps


class Kaiser(FIR):
    """Kaiser-window FIR (reference filtering/fir.py:52-137)."""

    def __init__(self, fpass, fstop, fs, gpass=1.0, gstop=40.0, **kwargs):
        super().__init__(fpass, fstop, gpass, gstop, fs, **kwargs)

    @property
    def _ripple(self):
        return max(self.pass_attenuation, self.gstop)

    @property
    def numtaps(self):
        ntaps, _ = sps.kaiserord(self._ripple, self.width / self.nyq)
        return _odd(ntaps)

    @property
    def window_params(self):
        return [sps.kaiser_beta(self._ripple)]


class _FixedWindow(FIR):
    """Windows whose attenuation is fixed: the width sets the tap count."""

    peak_err = None      # peak approximation error in dB
    width_factor = None  # taps = factor / (normalised transition width)

    def __init__(self, fpass, fstop, fs, **kwargs):
        gpass = -20 * np.log10(1 - 10 ** (self.peak_err / 20))
        super().__init__(fpass, fstop, gpass=gpass, gstop=self.peak_err, fs=fs, **kwargs)

    @property
    def numtaps(self):
        return _odd(int(self.width_factor / (self.width / self.nyq)))


class Rectangular(_FixedWindow):
    peak_err, width_factor = -21, 4

    @property
    def ftype(self):
        return "boxcar"


class Bartlett(_FixedWindow):
    peak_err, width_factor = -25, 8


class Hann(_FixedWindow):
    peak_err, width_factor = -44, 8


class Hamming(_FixedWindow):
    peak_err, width_factor = -53, 8


class Blackman(_FixedWindow):
    peak_err, width_factor = -74, 12


class Remez(FIR):
    """Parks-McClellan optimal FIR (reference filtering/fir.py:483-664)."""

    def __init__(self, bands, desired, fs, gpass=1, gstop=40, **kwargs):
        self.bands = np.array(bands).reshape(-1, 2)
        self.desired = np.array(desired, dtype=bool)
        fp = self.bands[self.desired].flatten()
        fpass = fp[np.logical_and(fp > 0, fp < fs / 2)]
        fst = self.bands[~self.desired].flatten()
        fstop = fst[np.logical_and(fst > 0, fst < fs / 2)]
        self.delta_pass = 1 - 10 ** (-gpass / 20)
        self.delta_stop = 10 ** (-gstop / 20)
        self.delta = self.delta_pass * self.desired + self.delta_stop * (1 - self.desired)
        super().__init__(fpass, fstop, gpass, gstop, fs, **kwargs)

    @property
    def btype(self):
        if len(self.fpass) < 2:
            return "lowpass" if self.fpass < self.fstop else "highpass"
        if len(self.fpass) == 2:
            return "bandstop" if self.fpass[0] < self.fstop[0] else "bandpass"
        return "multiband"

    @property
    def numtaps(self):
        """Bellanger's estimate."""
        n = -2 / 3 * np.log10(10 * self.delta_pass * self.delta_stop) * self.fs / self.width
        return _odd(int(np.ceil(n)))

    def _build(self, **kwargs):
        ntaps = kwargs.pop("numtaps", self.numtaps)
        weight = kwargs.pop("weight", 1 / self.delta)
        maxiter = kwargs.pop("maxiter", 25)
        grid_density = kwargs.pop("grid_density", 16)
        return sps.remez(ntaps, self.bands.flatten(), self.desired, weight=weight,
                         maxiter=maxiter, grid_density=grid_density, fs=self.fs, **kwargs)
