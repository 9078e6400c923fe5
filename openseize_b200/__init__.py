"""openseize_b200: a B200-native (sm_100a CUDA) implementation of openseize's
chunked filtering and spectral hot path behind the reference's producer /
operator API.

    from openseize_b200 import producer
    from openseize_b200.filtering.fir import Kaiser
    from openseize_b200.filtering.iir import Butter, Notch
    from openseize_b200.resampling.resampling import downsample, resample
    from openseize_b200.spectra.estimators import psd, stft

Importing the package does not touch CUDA; the first iteration of a producer
built by one of the operators does, and raises if no device or no built
library is present (there is no CPU fallback).
"""

from openseize_b200.core.producer import producer  # noqa: F401

__version__ = "0.1.0"


def set_compute(kind):
    """Arithmetic of the FFT-based FIR, the decimating polyphase filter and the
    Welch accumulation: "float64" (default, the reference's) or "float32" (opt-in:
    float64 samples in and out, the kernels' arithmetic in float32; results within
    ~1e-6 of the output peak, BASELINE's float32 tolerance is 1e-5).  IIR filters
    always run in float64.  Also settable with OSZ_COMPUTE=float32."""
    from openseize_b200.core import device

    device.set_compute(kind)


def set_io(kind):
    """Type of the samples in device memory and of the arrays handed back: "float64"
    (default: the reference returns float64 / complex128 for every input dtype) or
    "float32" (opt-in float32 I/O mode: float32 chunks cross PCIe and HBM as they are,
    results come back float32 / complex64, within 1e-5 of the output peak -- north_star's
    float32 tolerance; implies ``set_compute("float32")``; IIR recurrences, carried
    states and every sum stay float64).  Also settable with OSZ_IO=float32."""
    from openseize_b200.core import device

    device.set_io(kind)
