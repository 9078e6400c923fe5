"""Specialised filters (reference filtering/special.py): the Type III FIR
Hilbert transformer.  Design is host-side numpy/scipy; application is
``FIR.__call__`` -> GPU ``nm.oaconvolve`` like every other FIR."""

import numpy as np
import scipy.signal as sps

from openseize_b200.filtering.fir import Kaiser


class Hilbert(Kaiser):
    """Kaiser-windowed Type III (odd tap count, integer group delay) Hilbert
    transformer (reference filtering/special.py:16-133): ``x + 1j * Hilbert(x)``
    is the analytic signal.  Only the transition width is a parameter; the
    band always runs from ``width`` to ``nyquist - width``."""

    def __init__(self, width, fs, gpass=0.01, gstop=60):
        nyq = fs / 2
        super().__init__((0 + width, nyq - width), fstop=(0, nyq), fs=fs, gpass=gpass,
                         gstop=gstop)

    def _build(self, **kwargs):
        """Truncated ideal impulse response (1 - cos(pi n)) / (pi n), windowed
        (reference :120-133; Porat 1997, eqn 9.40)."""
        order = self.numtaps - 1
        taps = np.linspace(-order / 2, order / 2, self.numtaps)
        taps[order // 2] = 1                       # avoid 0/0 at the centre tap
        coeffs = (1 - np.cos(taps * np.pi)) / (taps * np.pi)
        coeffs[order // 2] = 0
        window = sps.get_window(("kaiser", *self.window_params), self.numtaps)
        return coeffs * window
