"""psd / stft with the reference's signatures and return values
(spectra/estimators.py:59-284)."""

import numpy as np

from openseize_b200.core import numerical as nm
from openseize_b200.core.producer import producer
from openseize_b200.core.resources import assignable


def psd(data, fs, axis=-1, resolution=0.5, window="hann", overlap=0.5, detrend="constant",
        scaling="density"):
    """Welch power spectrum (density) estimate.  Returns (number of averaged
    segments, frequencies, estimate).  The segment periodograms are averaged
    inside the GPU kernel instead of by the reference's host running mean
    (estimators.py:150-152); the result is the same mean."""
    pro = producer(data, chunksize=int(fs), axis=axis)        # estimators.py:141
    nfft = int(fs / resolution)                               # estimators.py:144
    freqs = np.fft.rfftfreq(nfft, 1 / fs)
    cnt, estimate = nm.welch_mean(pro, fs, nfft, window, overlap, axis, detrend, scaling)
    return cnt, freqs, estimate


def stft(data, fs, axis=-1, resolution=0.5, window="hann", overlap=0.5, detrend="constant",
         scaling="density", boundary=True, padded=True, asarray=True):
    """Short-time Fourier transform.  Returns (frequencies, segment times, X);
    X is an ndarray with the segments stacked on a new last axis when
    ``asarray`` and it fits in memory, else a producer of one complex array per
    segment."""
    pro = producer(data, chunksize=int(fs), axis=axis)
    nfft = int(fs / resolution)
    freqs, time, result = nm.stft(pro, fs, nfft, window, overlap, axis, detrend, scaling,
                                  boundary, padded)
    if asarray and assignable(result.shape):
        result = np.stack(list(result), axis=-1)
    return freqs, time, result
