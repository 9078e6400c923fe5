"""Parity cases shared by the CPU host-logic tests (numpy stand-in kernels,
`fake_gpu` fixture) and the GPU tests (real sm_100a kernels).

Each case drives the reference-facing operator API of openseize_b200 and
compares with (a) the golden vectors produced by the real reference
(tests/golden, made by oracle/make_golden.py) and (b) the oracle on fresh
seeded inputs.  Tolerance (BASELINE.json north_star): float64 results within
1e-9 of the output peak; lengths, counts, frequency and time vectors exact.
"""

import numpy as np

import oracle
from openseize_b200 import producer
from openseize_b200.core import numerical as nm
from openseize_b200.filtering.fir import Kaiser
from openseize_b200.filtering.iir import Butter, Notch
from openseize_b200.resampling.resampling import downsample, resample, upsample
from openseize_b200.spectra.estimators import psd, stft
from tests.conftest import golden, relerr, signal

TOL = 1e-9


def close(mine, ref, tol=TOL):
    mine, ref = np.asarray(mine), np.asarray(ref)
    assert mine.shape == ref.shape, (mine.shape, ref.shape)
    assert mine.dtype == ref.dtype, (mine.dtype, ref.dtype)
    err = relerr(mine, ref)
    assert err <= tol, err
    return err


# ------------------------------------------------------------------ FIR ----
def fir_golden():
    g = golden("fir_kaiser113")
    fs, cs = int(g["fs"]), int(g["chunksize"])
    x = signal(int(g["seed"]), int(g["rows"]), int(g["n"]), fs)
    assert x.sum() == float(g["x_sum"])
    filt = Kaiser(fpass=500, fstop=600, fs=fs)
    assert np.array_equal(filt.coeffs, g["taps"])
    for mode in ("same", "full", "valid"):
        y = filt(producer(x, cs, -1), cs, axis=-1, mode=mode).to_array()
        close(y, g["y_" + mode])
    # ndarray in -> ndarray out
    close(filt(x, cs, axis=-1, mode="same"), g["y_same"])
    # sample axis first
    xt = np.ascontiguousarray(x[:2].T)
    close(filt(producer(xt, cs, 0), cs, axis=0, mode="same").to_array(), g["y_same_axis0"])


def fir_long_golden():
    g = golden("fir_kaiser671")
    fs, cs = int(g["fs"]), int(g["chunksize"])
    x = signal(int(g["seed"]), int(g["rows"]), int(g["n"]), fs)
    filt = Kaiser(fpass=500, fstop=600, fs=fs)
    assert len(filt.coeffs) == 671 and np.array_equal(filt.coeffs, g["taps"])
    close(filt(producer(x, cs, -1), cs, axis=-1, mode="same").to_array(), g["y_same"])


def fir_oracle_sweep():
    """Reference tests/test_oaconvolve.py:14-94: lengths, axes, windows, modes."""
    import scipy.signal as sps

    rng = np.random.default_rng(0)
    for n in (10007, 33333):
        x = rng.standard_normal((3, 4, n))
        win = sps.get_window("hann", 203)
        for mode in ("same", "full", "valid"):
            got = np.concatenate(list(nm.oaconvolve(producer(x, 5000, -1), win, -1, mode)), -1)
            ref = np.concatenate(oracle.oaconvolve(x, win, 5000, -1, mode), -1)
            close(got, ref)
    x = rng.standard_normal((4, 6453, 2, 3))
    win = sps.get_window("blackman", 76)
    got = np.concatenate(list(nm.oaconvolve(producer(x, 1000, 1), win, 1, "same")), 1)
    close(got, np.concatenate(oracle.oaconvolve(x, win, 1000, 1, "same"), 1))
    x = rng.standard_normal((10622, 3))
    for k in (50, 77, 120):
        win = sps.get_window("hamming", k)
        got = np.concatenate(list(nm.oaconvolve(producer(x, 2048, 0), win, 0, "same")), 0)
        close(got, np.concatenate(oracle.oaconvolve(x, win, 2048, 0, "same"), 0))
    # short filters (direct-form kernel) and a chunk smaller than the filter
    x = rng.standard_normal((5, 9000))
    for k, cs in ((1, 1000), (2, 1000), (9, 333), (24, 4000), (25, 4000), (113, 100)):
        win = rng.standard_normal(k)
        for mode in ("same", "full", "valid"):
            got = np.concatenate(list(nm.oaconvolve(producer(x, cs, -1), win, -1, mode)), -1)
            ref = np.stack([np.convolve(r, win, mode) for r in x])
            close(got, ref)


def narrow_input_dtypes():
    """float32 / int16 chunks (EDF samples) travel narrow and come back float64
    (reference dtype rule)."""
    rng = np.random.default_rng(4)
    filt = Kaiser(fpass=500, fstop=600, fs=5000)
    for dtype in (np.float32, np.int16):
        x = (rng.standard_normal((3, 20000)) * 1000).astype(dtype)
        y = filt(producer(x, 6000, -1), 6000, axis=-1, mode="same").to_array()
        assert y.dtype == np.float64
        ref = np.concatenate(oracle.oaconvolve(x.astype(np.float64), filt.coeffs, 6000, -1,
                                               "same"), -1)
        close(y, ref)
        cnt, f, p = psd(x, 1024, resolution=1.0)
        rc, rf, rp = oracle.welch_psd(x.astype(np.float64), 1024, -1, 1.0)
        assert cnt == rc
        close(p, rp)


def float32_io():
    """Opt-in float32 I/O mode (north_star: 1e-5 of the signal peak in float32;
    reference dtype rule: SURVEY.md 8b, core/numerical.py:699): float32 chunks in,
    float32 / complex64 results out for FIR, IIR, resampling, Welch and STFT, against the
    float64 oracle.  The IIR recurrence stays float64 inside."""
    import openseize_b200

    rng = np.random.default_rng(41)
    fs, cs = 5000, 30000
    x64 = rng.standard_normal((3, 150000)) + 2.0
    x = x64.astype(np.float32)
    ref_in = x.astype(np.float64)
    kais = Kaiser(fpass=500, fstop=600, fs=fs)
    notch = Notch(fstop=60, width=6, fs=fs)
    butter = Butter(fpass=[1, 100], fstop=[0.5, 200], fs=fs, gpass=1, gstop=40)
    openseize_b200.set_io("float32")
    try:
        y = kais(producer(x, cs, -1), cs, axis=-1, mode="same").to_array()
        z = notch(producer(x, cs, -1), cs, axis=-1, dephase=True).to_array()
        zb = butter(producer(x, cs, -1), cs, axis=-1, dephase=True).to_array()
        d = downsample(producer(x, cs, -1), 4, fs, cs, axis=-1).to_array()
        cnt, f, p = psd(producer(x, cs, -1), fs, axis=-1, resolution=fs / 1024)
        fq, tq, X = stft(producer(x[:, :60000], cs, -1), fs, axis=-1, resolution=fs / 1024)
        # the chain on the device, float64 host input (narrowed on the host)
        chain = downsample(kais(notch(producer(x64, cs, -1), cs, axis=-1), cs, axis=-1), 4, fs, cs,
                           axis=-1)
        cnt2, _, p2 = psd(chain, fs / 4, axis=-1, resolution=fs / 4 / 256)
        # sample axis first
        yt = kais(producer(np.ascontiguousarray(x[:2].T), cs, 0), cs, axis=0).to_array()
    finally:
        openseize_b200.set_io("float64")
    for arr in (y, z, zb, d, p, yt, p2):
        assert arr.dtype == np.float32, arr.dtype
    assert X.dtype == np.complex64

    def within(got, ref, tol=1e-5):
        err = relerr(got.astype(ref.dtype), ref)
        assert 1e-12 < err <= tol, err          # float32 really was used, and it is close
        return err

    within(y, np.concatenate(oracle.oaconvolve(ref_in, kais.coeffs, cs, -1, "same"), -1))
    within(yt, np.concatenate(oracle.oaconvolve(ref_in[:2], kais.coeffs, cs, -1, "same"), -1).T)
    within(z, np.concatenate(oracle.filtfilt(ref_in, notch.coeffs, cs, -1), -1))
    within(zb, np.concatenate(oracle.sosfiltfilt(ref_in, butter.coeffs, cs, -1), -1))
    within(d, np.concatenate(oracle.polyphase_resample(ref_in, 1, 4, fs, cs, -1), -1))
    rc, rf, rp = oracle.welch_psd(ref_in, fs, -1, fs / 1024)
    assert cnt == rc and np.array_equal(f, rf)
    assert np.max(np.abs(p - rp) / np.max(rp, axis=-1, keepdims=True)) <= 1e-5
    of, ot, oX = oracle.stft(ref_in[:, :60000], fs, -1, fs / 1024)
    assert np.array_equal(fq, of) and np.array_equal(tq, ot)
    within(X, oX)
    r1 = np.concatenate(oracle.filtfilt(x64.astype(np.float32).astype(np.float64), notch.coeffs,
                                        cs, -1), -1)
    r2 = np.concatenate(oracle.oaconvolve(r1, kais.coeffs, cs, -1, "same"), -1)
    r3 = np.concatenate(oracle.polyphase_resample(r2, 1, 4, fs, cs, -1), -1)
    rc2, _, rp2 = oracle.welch_psd(r3, fs / 4, -1, fs / 4 / 256)
    assert cnt2 == rc2
    assert np.max(np.abs(p2 - rp2) / np.max(rp2, axis=-1, keepdims=True)) <= 1e-5
    # and the default mode is untouched afterwards
    y64 = kais(producer(x64, cs, -1), cs, axis=-1, mode="same").to_array()
    assert y64.dtype == np.float64


# ------------------------------------------------------------------ IIR ----
def iir_golden():
    g = golden("iir_butter8")
    fs = int(g["fs"])
    x = signal(int(g["seed"]), int(g["rows"]), int(g["n"]), fs)
    filt = Butter(fpass=[1, 100], fstop=[0.5, 200], fs=fs, gpass=1, gstop=40)
    assert np.array_equal(filt.coeffs, g["sos"])
    for cs in (4000, 5000):
        y = filt(producer(x, cs, -1), cs, axis=-1, dephase=True).to_array()
        close(y, g["y_filtfilt_cs%d" % cs])
    close(filt(producer(x, 4000, -1), 4000, axis=-1, dephase=False).to_array(), g["y_fwd_cs4000"])
    g = golden("iir_notch60")
    fs, cs = int(g["fs"]), int(g["chunksize"])
    x = signal(int(g["seed"]), int(g["rows"]), int(g["n"]), fs)
    notch = Notch(fstop=60, width=6, fs=fs)
    assert np.array_equal(notch.coeffs[0], g["b"]) and np.array_equal(notch.coeffs[1], g["a"])
    close(notch(producer(x, cs, -1), cs, axis=-1, dephase=True).to_array(), g["y_filtfilt"])
    close(notch(producer(x, cs, -1), cs, axis=-1, dephase=False).to_array(), g["y_fwd"])


def iir_oracle_sweep():
    """Reference tests/test_iir.py:77-158,216-326: axes, chunk sizes, zi."""
    import scipy.signal as sps

    rng = np.random.default_rng(9)
    sos = sps.butter(4, [1, 100], btype="bandpass", fs=5000, output="sos")
    # long chunks: several 8192-sample scan blocks per chunk, ragged tail
    x = rng.standard_normal((3, 70001)) + 3.0
    for cs in (70001, 30000, 8192, 8191, 1000):
        got = np.concatenate(list(nm.sosfiltfilt(producer(x, cs, -1), sos, -1)), -1)
        close(got, np.concatenate(oracle.sosfiltfilt(x, sos, cs, -1), -1))
        got = np.concatenate(list(nm.sosfilt(producer(x, cs, -1), sos, -1)), -1)
        close(got, np.concatenate(oracle.sosfilt(x, sos, cs, -1)[0], -1))
    # sample axis in the middle, explicit zi
    x = rng.standard_normal((2, 20011, 3))
    zi = rng.standard_normal((sos.shape[0], 2, 2, 3))
    got = np.concatenate(list(nm.sosfilt(producer(x, 7000, 1), sos, 1, zi)), 1)
    close(got, np.concatenate(oracle.sosfilt(x, sos, 7000, 1, zi)[0], 1))
    got = np.concatenate(list(nm.sosfiltfilt(producer(x, 7000, 1), sos, 1)), 1)
    close(got, np.concatenate(oracle.sosfiltfilt(x, sos, 7000, 1), 1))
    # transfer-function format (Notch path), forward with zi and forward-backward
    b, a = sps.iirnotch(60, 10, fs=5000)
    x = rng.standard_normal((4, 25000))
    zi = rng.standard_normal((4, 2))
    got = np.concatenate(list(nm.lfilter(producer(x, 6000, -1), (b, a), -1, zi)), -1)
    close(got, np.concatenate(oracle.lfilter(x, (b, a), 6000, -1, zi)[0], -1))
    got = np.concatenate(list(nm.filtfilt(producer(x, 6000, -1), (b, a), -1)), -1)
    close(got, np.concatenate(oracle.filtfilt(x, (b, a), 6000, -1), -1))
    # transfer-function filters above second order (sequential DF2T kernel)
    for coeffs in (sps.butter(4, [0.05, 0.3], btype="bandpass"), sps.cheby1(5, 1, 0.25),
                   sps.butter(3, 0.2)):
        k = max(len(coeffs[0]), len(coeffs[1])) - 1
        x = rng.standard_normal((3, 21013)) + 1.0
        zi = rng.standard_normal((3, k))
        got = np.concatenate(list(nm.lfilter(producer(x, 6000, -1), coeffs, -1, zi)), -1)
        close(got, np.concatenate(oracle.lfilter(x, coeffs, 6000, -1, zi)[0], -1))
        got = np.concatenate(list(nm.filtfilt(producer(x, 6000, -1), coeffs, -1)), -1)
        close(got, np.concatenate(oracle.filtfilt(x, coeffs, 6000, -1), -1))
    xm = rng.standard_normal((2, 9001, 2))
    coeffs = sps.butter(4, 0.3)
    got = np.concatenate(list(nm.filtfilt(producer(xm, 4000, 1), coeffs, 1)), 1)
    close(got, np.concatenate(oracle.filtfilt(xm, coeffs, 4000, 1), 1))
    # chunks longer than the cascade's settle length: the look-ahead pass of the
    # forward-backward filters is cut to the samples that still matter
    x = rng.standard_normal((2, 130001)) + 5.0
    from openseize_b200.core.numerical import _ba_to_sos

    settle = nm._Cascade(_ba_to_sos((b, a))).settle
    assert settle < 40000
    got = np.concatenate(list(nm.filtfilt(producer(x, 40000, -1), (b, a), -1)), -1)
    close(got, np.concatenate(oracle.filtfilt(x, (b, a), 40000, -1), -1), tol=1e-12)
    lp = sps.butter(4, 300, fs=5000, output="sos")
    assert nm._Cascade(lp).settle < 40000
    got = np.concatenate(list(nm.sosfiltfilt(producer(x, 40000, -1), lp, -1)), -1)
    close(got, np.concatenate(oracle.sosfiltfilt(x, lp, 40000, -1), -1), tol=1e-12)
    # a cascade of eight IDENTICAL sections: transients of coinciding poles decay like
    # n^7 r^n -- the settle length has to account for the multiplicity
    one = sps.butter(1, [40, 60], btype="bandpass", fs=5000, output="sos")
    rep = np.repeat(one, 8, axis=0)
    s1, s8 = nm._Cascade(one).settle, nm._Cascade(rep).settle
    assert s1 < s8 < 20000
    xr = rng.standard_normal((2, 90001)) + 2.0
    got = np.concatenate(list(nm.sosfiltfilt(producer(xr, 30000, -1), rep, -1)), -1)
    close(got, np.concatenate(oracle.sosfiltfilt(xr, rep, 30000, -1), -1), tol=1e-12)
    # more than 16 sections: split cascade
    sos = sps.butter(20, [5, 400], btype="bandpass", fs=5000, output="sos")
    assert sos.shape[0] == 20
    x = rng.standard_normal((2, 30000))
    got = np.concatenate(list(nm.sosfilt(producer(x, 9000, -1), sos, -1)), -1)
    close(got, np.concatenate(oracle.sosfilt(x, sos, 9000, -1)[0], -1))


# ------------------------------------------------------------- resample ----
def resample_golden():
    g = golden("resample")
    fs = int(g["fs"])
    xfull = signal(int(g["seed"]), int(g["rows"]), int(g["n"]), fs)
    for name in ("down20", "up2", "rs3_7"):
        L, M, cs, nuse = (int(v) for v in g["LMcsn_" + name])
        x = xfull[:, :nuse]
        pro = resample(producer(x, cs, -1), L, M, fs, cs, axis=-1)
        assert pro.shape == g["y_" + name].shape
        close(pro.to_array(), g["y_" + name])
        # raw generator: bit-exact yield lengths (reference numerical.py:617-632)
        raw = [a.shape[-1] for a in
               nm.polyphase_resample(producer(x, cs, -1), L, M, fs, Kaiser, -1)]
        assert raw == list(g["raw_" + name]), (raw, list(g["raw_" + name]))
    x = xfull[:, :30000]
    close(downsample(x, 20, fs, 7000, axis=-1), g["y_down20"])
    close(upsample(xfull[:, :9000], 2, fs, 2500, axis=-1), g["y_up2"])


def resample_oracle_sweep():
    """Reference tests/test_resampling.py:39-139: all L != M in 1..5, chunk
    sizes, sizes x channels; the oracle is one global resample_poly call."""
    import scipy.signal as sps

    rng = np.random.default_rng(33)
    x = rng.standard_normal((6, 39968))
    fs = 500
    for L in range(1, 6):
        for M in range(1, 6):
            if np.gcd(L, M) != 1 or L == M:
                continue
            got = resample(producer(x, 10000, -1), L, M, fs, 10000, axis=-1).to_array()
            h = oracle.resample_filter(L, M, fs)
            ref = sps.resample_poly(x, L, M, axis=-1, window=h)
            close(got, ref)
    for cs in (1000, 3333, 13000, 39968):
        got = downsample(producer(x, cs, -1), 10, fs, cs, axis=-1).to_array()
        ref = sps.resample_poly(x, 1, 10, axis=-1, window=oracle.resample_filter(1, 10, fs))
        close(got, ref)
    xt = np.ascontiguousarray(x[:3].T)
    got = downsample(producer(xt, 9000, 0), 4, fs, 9000, axis=0).to_array()
    ref = sps.resample_poly(xt, 1, 4, axis=0, window=oracle.resample_filter(1, 4, fs))
    close(got, ref)
    assert resample(x, 3, 3, fs, 1000) is x          # identity returns the input itself
    # non-coprime factors straight into nm.polyphase_resample (the public resample()
    # reduces them first, resampling.py; scipy's resample_poly reduces them itself)
    from openseize_b200.filtering.fir import Kaiser as _K

    for L, M in ((2, 4), (4, 6), (6, 4)):
        got = list(nm.polyphase_resample(producer(x, 10000, -1), L, M, fs, _K, -1))
        ref = oracle.polyphase_resample(x, L, M, fs, 10000, -1)
        assert [a.shape[-1] for a in got] == [a.shape[-1] for a in ref]
        close(np.concatenate(got, -1), np.concatenate(ref, -1))


# -------------------------------------------------------------- spectra ----
def spectra_golden(name="pow2"):
    g = golden("spectra_" + name)
    fs, res = int(g["fs"]), float(g["resolution"])
    x = signal(int(g["seed"]), int(g["rows"]), int(g["n"]), fs)
    for det in ("constant", "linear"):
        for scal in ("density", "spectrum"):
            cnt, f, p = psd(producer(x, 5000, -1), fs, axis=-1, resolution=res, detrend=det,
                            scaling=scal)
            assert cnt == int(g["psd_cnt"]) and np.array_equal(f, g["freqs"])
            close(p, g["psd_%s_%s" % (det, scal)])
    for bnd in (True, False):
        for pad in (True, False):
            f, t, X = stft(producer(x, 5000, -1), fs, axis=-1, resolution=res, boundary=bnd,
                           padded=pad)
            key = "stft_b%d_p%d" % (bnd, pad)
            assert np.array_equal(f, g["freqs"]) and np.array_equal(t, g[key + "_time"])
            assert X.shape[-1] == int(g[key + "_nseg"]) and X.dtype == np.complex128
            close(X[..., g[key + "_idx"]], g[key + "_X"])


def spectra_oracle_sweep(fs=1024, resolutions=(1.0, 2.0, 4.0)):
    """Reference tests/test_spectra.py:164-615: sizes, overlaps, windows,
    scalings, axes; welch producer and stft producer paths."""
    rng = np.random.default_rng(1234)
    x = rng.standard_normal((3, 41017)) + 2.0
    for res in resolutions:
        for ov in (0.1, 0.5, 0.75):
            for win in ("hann", "hamming"):
                cnt, f, p = psd(producer(x, 7000, -1), fs, resolution=res, window=win, overlap=ov)
                rc, rf, rp = oracle.welch_psd(x, fs, -1, res, window=win, overlap=ov)
                assert cnt == rc and np.array_equal(f, rf)
                close(p, rp)
    # sample axis first; ndarray input
    xt = np.ascontiguousarray(x[:2, :20000].T)
    cnt, f, p = psd(xt, fs, axis=0, resolution=resolutions[0])
    rc, rf, rp = oracle.welch_psd(xt, fs, 0, resolutions[0])
    assert cnt == rc
    close(p, rp)
    # welch as a producer of per-segment periodograms
    nfft = int(fs / resolutions[0])
    f, wpro = nm.welch(producer(x, 6000, -1), fs, nfft, "hann", 0.5, -1, "constant", "density")
    segs = oracle.segments(x, nfft, 0.5, 6000, -1)
    got = list(wpro)
    assert len(got) == len(segs) == wpro.shape[-1]
    for k in (0, 1, len(segs) - 1):
        close(got[k], oracle.periodogram(segs[k], fs, nfft)[1])
    # stft, all boundary / padding combinations, producer output
    for bnd in (True, False):
        for pad in (True, False):
            f, t, X = stft(producer(x, 9000, -1), fs, resolution=resolutions[0], boundary=bnd,
                           padded=pad, overlap=0.5)
            rf, rt, rX = oracle.stft(x, fs, -1, resolutions[0], boundary=bnd, padded=pad)
            assert np.array_equal(t, rt)
            close(X, rX)
    f, t, Xp = stft(producer(x, 9000, -1), fs, resolution=resolutions[0], asarray=False)
    assert not isinstance(Xp, np.ndarray)
    close(np.stack(list(Xp), -1), oracle.stft(x, fs, -1, resolutions[0])[2])
    # single in-memory segments, zero padded (nfft > samples) and cropped
    # (reference tests/test_spectra.py:16-139)
    seg = rng.standard_normal((3, 700)) + 1.5
    for nfft in (700, 1024, 1500, 512):
        for det in ("constant", "linear"):
            for scaling in ("density", "spectrum"):
                f, p = nm.periodogram(seg, fs, nfft=nfft, window="hann", axis=-1, detrend=det,
                                      scaling=scaling)
                rf, rp = oracle.periodogram(seg, fs, nfft, "hann", -1, det, scaling)
                assert np.array_equal(f, rf)
                close(p, rp)
    f, X = nm.modified_dft(seg.T.copy(), fs, 1024, "hamming", 0, "constant", "density")
    rf, rX = oracle.modified_dft(seg.T.copy(), fs, 1024, "hamming", 0, "constant", "density")
    close(X, rX)


def fused_fir_decimate():
    """downsample(FIR(x)) is run as one decimating filter with the edges done by
    the two unfused kernels: it must equal the reference's two stages on every
    sample, for ragged lengths, chunk sizes and factors; OSZ_FUSE=0 (unfused
    device chain) gives the same."""
    import os

    rng = np.random.default_rng(12)
    fs = 5000
    kais = Kaiser(fpass=500, fstop=600, fs=fs)                 # 113 taps
    for n, cs, M in ((30011, 7000, 4), (52345, 10000, 10), (20000, 3333, 3), (41000, 41000, 5)):
        x = rng.standard_normal((3, n)) + 2.0
        r1 = np.concatenate(oracle.oaconvolve(x, kais.coeffs, cs, -1, "same"), -1)
        ref_blocks = oracle.polyphase_resample(r1, 1, M, fs, cs, -1)
        ref = np.concatenate(ref_blocks, -1)
        for fuse in ("1", "0"):
            os.environ["OSZ_FUSE"] = fuse
            try:
                pro = downsample(kais(producer(x, cs, -1), cs, axis=-1), M, fs, cs, axis=-1)
                got_blocks = [np.array(b) for b in pro]
            finally:
                os.environ.pop("OSZ_FUSE", None)
            got = np.concatenate(got_blocks, -1)
            assert got.shape == ref.shape == pro.shape, (fuse, got.shape, ref.shape)
            close(got, ref, tol=1e-12)
            # the edges are where the fused filter alone would be wrong
            close(got[:, :200], ref[:, :200], tol=1e-12)
            close(got[:, -200:], ref[:, -200:], tol=1e-12)
    # not fusable (mode 'valid'): plain device chain
    x = rng.standard_normal((2, 30000))
    r1 = np.concatenate(oracle.oaconvolve(x, kais.coeffs, 7000, -1, "valid"), -1)
    ref = np.concatenate(oracle.polyphase_resample(r1, 1, 4, fs, 7000, -1), -1)
    got = downsample(kais(producer(x, 7000, -1), 7000, axis=-1, mode="valid"), 4, fs, 7000,
                     axis=-1).to_array()
    close(got, ref, tol=1e-12)


def fused_iir_fir_decimate():
    """downsample(FIR_same(IIR_dephase(x))) runs the IIR's backward pass, the FIR and
    the decimator as one kernel per chunk (outputs straddling chunk / span boundaries
    from the exported edge samples, the recording's ends by the unfused kernels): it
    must equal the reference's three stages on every sample -- ragged lengths, chunk
    sizes around the filter length, both IIR formats; OSZ_FUSE_IIR=0 gives the same."""
    import os

    import scipy.signal as sps

    rng = np.random.default_rng(21)
    fs = 5000
    kais = Kaiser(fpass=500, fstop=600, fs=fs)                 # 113 taps
    notch = Notch(fstop=60, width=6, fs=fs)
    butter = Butter(fpass=[5, 400], fstop=[2, 600], fs=fs, gpass=1, gstop=30)
    cases = ((notch, 60011, 9000, 4), (notch, 52345, 20000, 10), (butter, 41000, 6500, 3),
             (notch, 30000, 30000, 5), (butter, 25013, 2500, 4), (notch, 15000, 700, 2))
    for filt, n, cs, M in cases:
        x = rng.standard_normal((3, n)) + 2.0
        if filt is notch:
            r0 = np.concatenate(oracle.filtfilt(x, filt.coeffs, cs, -1), -1)
        else:
            r0 = np.concatenate(oracle.sosfiltfilt(x, filt.coeffs, cs, -1), -1)
        r1 = np.concatenate(oracle.oaconvolve(r0, kais.coeffs, cs, -1, "same"), -1)
        ref = np.concatenate(oracle.polyphase_resample(r1, 1, M, fs, cs, -1), -1)
        rc, rf, rp = oracle.welch_psd(ref, fs / M, -1, fs / M / 256)
        for fuse in ("1", "0"):
            os.environ["OSZ_FUSE_IIR"] = fuse
            before = dict(nm.FUSED_STATS)
            try:
                def chain():
                    p1 = filt(producer(x, cs, -1), cs, axis=-1, dephase=True)
                    return downsample(kais(p1, cs, axis=-1), M, fs, cs, axis=-1)

                # a device consumer (psd) takes the fused path; a host consumer the chunk grid
                cnt, f, p = psd(chain(), fs / M, axis=-1, resolution=fs / M / 256)
                got = chain().to_array()
            finally:
                os.environ.pop("OSZ_FUSE_IIR", None)
            used = nm.FUSED_STATS["fused_chunks"] + nm.FUSED_STATS["fallback_chunks"] \
                - before["fused_chunks"] - before["fallback_chunks"]
            assert (used > 0) == (fuse == "1"), (fuse, used)
            assert cnt == rc and np.array_equal(f, rf)
            close(p, rp)
            close(got, ref, tol=1e-11)
    # the decimated stream itself through the fused path (device consumer = a GPU stage
    # that does not care about the block grid): every sample, edges included
    x = rng.standard_normal((2, 47001)) + 1.0
    cs, M = 9000, 4
    r0 = np.concatenate(oracle.filtfilt(x, notch.coeffs, cs, -1), -1)
    r1 = np.concatenate(oracle.oaconvolve(r0, kais.coeffs, cs, -1, "same"), -1)
    ref = np.concatenate(oracle.polyphase_resample(r1, 1, M, fs, cs, -1), -1)
    pro = downsample(kais(notch(producer(x, cs, -1), cs, axis=-1), cs, axis=-1), M, fs, cs, axis=-1)
    twin = nm._device_twin(pro)
    blocks = twin[0](*twin[1], **dict(twin[2], _free=True))
    got = np.concatenate(list(nm._to_host(blocks, nm._layout_of(pro, -1))), -1)
    close(got, ref, tol=1e-11)
    close(got[:, :100], ref[:, :100], tol=1e-11)
    close(got[:, -100:], ref[:, -100:], tol=1e-11)


# ------------------------------------------------------------- pipeline ----
def pipeline_chain():
    """Notch -> Kaiser FIR -> downsample -> psd, composed through producers
    (config 5 of BASELINE.json at toy size).  On this package the chain stays on
    the device; the oracle composes the same stages on whole arrays with the
    chunk sizes the reference would have used."""
    fs, cs = 5000, 8000
    x = signal(77, 3, 60000, fs)
    notch = Notch(fstop=60, width=6, fs=fs)
    kais = Kaiser(fpass=500, fstop=600, fs=fs)
    p1 = notch(producer(x, cs, -1), cs, axis=-1, dephase=True)
    p2 = kais(p1, cs, axis=-1, mode="same")
    p3 = downsample(p2, 4, fs, cs, axis=-1)
    cnt, f, p = psd(p3, fs // 4, axis=-1, resolution=fs / 4 / 256)

    r1 = np.concatenate(oracle.filtfilt(x, notch.coeffs, cs, -1), -1)
    r2 = np.concatenate(oracle.oaconvolve(r1, kais.coeffs, cs, -1, "same"), -1)
    r3 = np.concatenate(oracle.polyphase_resample(r2, 1, 4, fs, cs, -1), -1)
    rc, rf, rp = oracle.welch_psd(r3, fs // 4, -1, fs / 4 / 256)
    assert cnt == rc and np.array_equal(f, rf)
    close(p, rp)
    # and the intermediate producer is still iterable on the host
    close(kais(notch(producer(x, cs, -1), cs, axis=-1), cs, axis=-1).to_array(), r2)


def c5_real_parameters(rows=2, nchunks=6, tail=123_457, cs=1_000_000):
    """BASELINE.json config 5 -- the configuration bench.py measures -- with its
    REAL parameters: Notch(60 Hz, width 6) at 30 kHz forward-backward (the
    settle-shortened look-ahead is active: settle ~ 88 k < chunksize 1e6) ->
    Kaiser(500, 600) at 30 kHz (671 taps, 'same') -> downsample M = 25 (561-tap
    anti-alias; fused: 1231 taps) -> psd at 1200 Hz with nfft 4096; chunksize
    1e6, a ragged last chunk.  Every stage's output against the oracle at
    1e-9 of its peak, Welch count and frequencies exact, with the stage fusion
    on and off.  Reference: core/numerical.py:449-520,158-298,523-632,852-947."""
    import os

    fs, M, nfft = 30000, 25, 4096
    n = nchunks * cs + tail
    x = signal(5, rows, n, fs)
    notch = Notch(fstop=60, width=6, fs=fs)
    kais = Kaiser(fpass=500, fstop=600, fs=fs)
    assert len(kais.coeffs) == 671
    r1 = np.concatenate(oracle.filtfilt(x, notch.coeffs, cs, -1), -1)
    r2 = np.concatenate(oracle.oaconvolve(r1, kais.coeffs, cs, -1, "same"), -1)
    r3_blocks = oracle.polyphase_resample(r2, 1, M, fs, cs, -1)
    r3 = np.concatenate(r3_blocks, -1)
    fs2 = fs // M
    rc, rf, rp = oracle.welch_psd(r3, fs2, -1, fs2 / nfft)
    assert rc == (r3.shape[-1] - nfft) // (nfft // 2) + 1
    errs = {}
    for fuse, fuse_iir in (("1", "0"), ("0", "0"), ("1", "1")):
        os.environ["OSZ_FUSE"] = fuse
        os.environ["OSZ_FUSE_IIR"] = fuse_iir
        try:
            def chain():
                p1 = notch(producer(x, cs, -1), cs, axis=-1, dephase=True)
                p2 = kais(p1, cs, axis=-1, mode="same")
                return downsample(p2, M, fs, cs, axis=-1)

            cnt, f, p = psd(chain(), fs2, axis=-1, resolution=fs2 / nfft)
            d = chain()
            got = [np.array(b) for b in d]
        finally:
            os.environ.pop("OSZ_FUSE", None)
            os.environ.pop("OSZ_FUSE_IIR", None)
        assert cnt == rc and np.array_equal(f, rf), (fuse, cnt, rc)
        assert d.shape == r3.shape
        fuse = fuse + fuse_iir
        errs["decimated_fuse" + fuse] = close(np.concatenate(got, -1), r3)
        # per-bin error relative to each channel's largest bin (SURVEY 8d parity metric)
        e = float(np.max(np.abs(p - rp) / np.max(rp, axis=-1, keepdims=True)))
        assert e <= TOL, (fuse, e)
        errs["psd_fuse" + fuse] = e
    # the full-rate intermediate stages on their own (host-visible producers)
    y1 = notch(producer(x, cs, -1), cs, axis=-1, dephase=True).to_array()
    errs["notch"] = close(y1, r1)
    y2 = kais(notch(producer(x, cs, -1), cs, axis=-1, dephase=True), cs, axis=-1).to_array()
    errs["fir"] = close(y2, r2)
    return errs


def hilbert_golden():
    """Type III Hilbert transformer (reference filtering/special.py): design
    bit-identical, application within tolerance of the reference's output."""
    from openseize_b200.filtering.special import Hilbert

    g = golden("hilbert")
    fs, cs = int(g["fs"]), int(g["chunksize"])
    x = signal(int(g["seed"]), int(g["rows"]), int(g["n"]), fs)
    hil = Hilbert(width=float(g["width"]), fs=fs)
    assert np.array_equal(hil.coeffs, g["taps"])
    close(hil(producer(x, cs, -1), cs, axis=-1).to_array(), g["y"])


def analytic_golden():
    """Analytic signal x + i H(x) (reference experimental/coupling/transforms.py:107-192):
    complex signal, amplitudes and phases against the real reference's outputs."""
    from openseize_b200.experimental.coupling.transforms import Analytic

    g = golden("analytic")
    fs, cs = int(g["fs"]), int(g["chunksize"])
    x = signal(int(g["seed"]), int(g["rows"]), int(g["n"]), fs)
    assert x.sum() == float(g["x_sum"])
    ana = Analytic(x, fs, chunksize=cs, axis=-1, width=float(g["width"]))
    z = ana.signal.to_array()
    assert z.dtype == np.complex128 and ana.signal.shape == x.shape
    close(z, g["z"])
    close(ana.amplitudes.to_array(), g["amplitudes"])
    ph, rph = ana.phases.to_array(), g["phases"]
    d = np.abs(ph - rph)
    assert np.max(np.minimum(d, 2 * np.pi - d)) < 1e-9          # phases wrap at 2 pi
    # sample axis first, chunk size that does not divide the recording
    xt = np.ascontiguousarray(x.T)
    zt = Analytic(xt, fs, chunksize=1777, axis=0, width=float(g["width"])).signal.to_array()
    close(zt, g["z"].T)


# ------------------------------------------------- producer tools (N3) ----
def protools_golden():
    """Masked producers consumed by GPU operators and protools.mean / std /
    standardize against the real reference's outputs (tests/golden/protools.npz)."""
    from openseize_b200.core import protools

    g = golden("protools")
    fs, cs, mask = int(g["fs"]), int(g["chunksize"]), g["mask"]
    x = signal(int(g["seed"]), int(g["rows"]), int(g["n"]), fs) + float(g["offset"])
    assert x.sum() == float(g["x_sum"])
    for name, arr, axis in (("ax1", x, -1), ("ax0", np.ascontiguousarray(x.T), 0)):
        # host iteration of the masked producer: the reference's yields, bit for bit
        mpro = producer(arr, cs, axis, mask=mask)
        got = list(mpro)
        assert [a.shape[axis] for a in got] == list(g["masked_lengths_" + name])
        assert np.array_equal(np.concatenate(got, axis), g["masked_" + name])
        assert mpro.shape[axis] == g["masked_" + name].shape[axis]
        # the device compaction (what a GPU operator downstream consumes): same samples
        layout_axis = axis % arr.ndim
        blocks = list(nm._to_host(nm.device_chunks(producer(arr, cs, axis, mask=mask), axis),
                                  nm._layout_of(mpro, layout_axis)))
        assert [a.shape[axis] for a in blocks] == list(g["masked_lengths_" + name])
        assert np.array_equal(np.concatenate(blocks, axis), g["masked_" + name])
        for ax in (0, 1):
            for keep in (0, 1):
                m = protools.mean(producer(arr, cs, axis), ax, keepdims=bool(keep))
                s = protools.std(producer(arr, cs, axis), ax, keepdims=bool(keep))
                close(m, g["mean_%s_axis%d_keep%d" % (name, ax, keep)], 1e-12)
                close(s, g["std_%s_axis%d_keep%d" % (name, ax, keep)], 1e-9)
            z = protools.standardize(producer(arr, cs, axis), ax)
            assert z.shape == arr.shape and z.chunksize == cs
            close(z.to_array(), g["standardized_%s_axis%d" % (name, ax)], 1e-9)
    # NaNs are skipped per chunk; a chunk of nothing but NaNs makes the row NaN
    # (the reference's n * nanmean(chunk) weighting, protools.py:531-536)
    xn = x.copy()
    for r, a, b in g["nan_spans"]:
        xn[r, a:b] = np.nan
    m = protools.mean(producer(xn, cs, -1), -1)
    s = protools.std(producer(xn, cs, -1), -1)
    assert np.array_equal(np.isnan(m), np.isnan(g["mean_nan"]))
    ok = ~np.isnan(m)
    close(m[ok], g["mean_nan"][ok], 1e-12)
    close(s[ok], g["std_nan"][ok], 1e-9)


def masked_chain():
    """The quickstart workflow: filter -> state mask ->
    standardize -> PSD, every stage handing its chunks over on the device;
    against the oracle run stage by stage."""
    from openseize_b200.core import protools

    fs, cs = 1000, 4000
    rng = np.random.default_rng(61)
    x = signal(62, 3, 30000, fs) + 7.0
    mask = np.repeat(rng.random(60) < 0.6, 500)
    filt = Notch(fstop=60, width=6, fs=fs)
    y = filt(producer(x, cs, -1), cs, axis=-1, dephase=False)
    masked = producer(y, cs, -1, mask=mask)
    z = protools.standardize(masked, -1)
    cnt, freqs, est = psd(z, fs, resolution=fs / 1024)
    b, a = filt.coeffs
    oy = np.concatenate(oracle.lfilter(x, (b, a), cs, -1)[0], -1)
    om = oracle.masked(oy, mask, cs, -1)
    oz = np.concatenate(oracle.standardize(om, -1, -1), -1)
    ocnt, ofreqs, oest = oracle.welch_psd(oz, fs, -1, fs / 1024)
    assert cnt == ocnt and np.array_equal(freqs, ofreqs)
    close(est, oest)
    close(z.to_array(), oz)


def protools_edges():
    """Masks shorter than the data, masked producers straight into psd and into a
    FIR, scalar results for 1-D producers, 3-D producers, pickling."""
    import pickle

    from openseize_b200.core import protools

    rng = np.random.default_rng(0)
    x = rng.standard_normal((3, 20000)) + 5
    mask = rng.random(12000) < 0.5                      # production stops with the mask
    ref = np.concatenate(oracle.masked(x, mask, 3000, -1), -1)
    mpro = producer(x, 3000, -1, mask=mask)
    got = np.concatenate(list(nm._to_host(nm.device_chunks(mpro, -1), nm._layout_of(mpro, -1))), -1)
    assert np.array_equal(ref, got)
    cnt, _, p = psd(producer(x, 3000, -1, mask=mask), 1000.0, resolution=1000 / 512)
    ocnt, _, op = oracle.welch_psd(ref, 1000.0, -1, 1000 / 512)
    assert cnt == ocnt
    close(p, op)
    filt = Kaiser(100, 150, 1000)
    y = filt(producer(x, 3000, -1, mask=mask), 3000, axis=-1).to_array()
    close(y, np.concatenate(oracle.oaconvolve(ref, filt.coeffs, 3000, -1, "same"), -1))
    m = protools.mean(producer(x[0], 3000, -1), -1)
    s = protools.std(producer(x[0], 3000, -1), -1)
    assert np.ndim(m) == 0 and abs(m - x[0].mean()) < 1e-12 and abs(s - x[0].std()) < 1e-10
    x3 = rng.standard_normal((2, 3, 7000))
    m3 = protools.mean(producer(x3, 2000, -1), -1)
    assert m3.shape == (2, 3)
    close(m3, x3.mean(-1), 1e-12)
    z3 = protools.standardize(producer(x3, 2000, -1), -1).to_array()
    close(z3, (x3 - x3.mean(-1, keepdims=True)) / x3.std(-1, keepdims=True))
    zp = protools.standardize(producer(x, 3000, -1), -1)
    assert np.array_equal(pickle.loads(pickle.dumps(zp)).to_array(), zp.to_array())
