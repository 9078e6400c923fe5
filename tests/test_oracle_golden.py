"""The oracle against the golden vectors written by the REAL reference
(oracle/make_golden.py).  Runs everywhere; needs no GPU and no /root/reference."""

import numpy as np
import scipy.signal as sps

import oracle
from tests.conftest import golden, signal


def _x(g):
    x = signal(int(g["seed"]), int(g["rows"]), int(g["n"]), int(g["fs"]))
    assert x.sum() == float(g["x_sum"]), "seeded input differs from the one the golden was made on"
    return x


def test_fir_golden_bit_exact():
    g = golden("fir_kaiser113")
    x, cs = _x(g), int(g["chunksize"])
    for mode in ("same", "full", "valid"):
        blocks = oracle.oaconvolve(x, g["taps"], cs, -1, mode)
        assert np.array_equal(np.concatenate(blocks, -1), g["y_" + mode])
    blocks = oracle.oaconvolve(x, g["taps"], cs, -1, "same")
    assert [b.shape[-1] for b in blocks] == list(g["raw_block_lengths_same"])
    xt = np.ascontiguousarray(x[:2].T)
    assert np.array_equal(np.concatenate(oracle.oaconvolve(xt, g["taps"], cs, 0, "same"), 0),
                          g["y_same_axis0"])
    g = golden("fir_kaiser671")
    x = _x(g)
    got = np.concatenate(oracle.oaconvolve(x, g["taps"], int(g["chunksize"]), -1, "same"), -1)
    assert np.array_equal(got, g["y_same"])


def test_iir_golden_bit_exact():
    g = golden("iir_butter8")
    x = _x(g)
    for cs in (4000, 5000):
        got = np.concatenate(oracle.sosfiltfilt(x, g["sos"], cs, -1), -1)
        assert np.array_equal(got, g["y_filtfilt_cs%d" % cs])
    assert np.array_equal(np.concatenate(oracle.sosfilt(x, g["sos"], 4000, -1)[0], -1),
                          g["y_fwd_cs4000"])
    g = golden("iir_notch60")
    x, cs = _x(g), int(g["chunksize"])
    assert np.array_equal(np.concatenate(oracle.filtfilt(x, (g["b"], g["a"]), cs, -1), -1),
                          g["y_filtfilt"])
    assert np.array_equal(np.concatenate(oracle.lfilter(x, (g["b"], g["a"]), cs, -1)[0], -1),
                          g["y_fwd"])


def test_resample_golden_bit_exact():
    g = golden("resample")
    xfull = _x(g)
    for name in ("down20", "up2", "rs3_7"):
        L, M, cs, nuse = (int(v) for v in g["LMcsn_" + name])
        x = xfull[:, :nuse]
        blocks = oracle.polyphase_resample(x, L, M, int(g["fs"]), cs, -1)
        assert [b.shape[-1] for b in blocks] == list(g["raw_" + name])
        assert np.array_equal(np.concatenate(blocks, -1), g["y_" + name])
        assert np.array_equal(oracle.resample_filter(L, M, int(g["fs"])), g["h_" + name])
        # the chunked algorithm equals ONE global resample_poly call
        assert np.array_equal(sps.resample_poly(x, L, M, axis=-1, window=g["h_" + name]),
                              g["y_" + name])


def test_spectra_golden_bit_exact():
    for name in ("pow2", "nonpow2"):
        g = golden("spectra_" + name)
        x, fs, res = _x(g), int(g["fs"]), float(g["resolution"])
        for det in ("constant", "linear"):
            for scal in ("density", "spectrum"):
                cnt, f, p = oracle.welch_psd(x, fs, -1, res, detrend=det, scaling=scal)
                assert cnt == int(g["psd_cnt"]) and np.array_equal(f, g["freqs"])
                assert np.array_equal(p, g["psd_%s_%s" % (det, scal)])
        for bnd in (True, False):
            for pad in (True, False):
                f, t, X = oracle.stft(x, fs, -1, res, boundary=bnd, padded=pad)
                key = "stft_b%d_p%d" % (bnd, pad)
                assert np.array_equal(t, g[key + "_time"]) and X.shape[-1] == int(g[key + "_nseg"])
                assert np.array_equal(X[..., g[key + "_idx"]], g[key + "_X"])


def test_oracle_matches_scipy_global_calls():
    """The reference's own acceptance tests compare with scipy's in-memory
    routines (tests/test_oaconvolve.py, test_iir.py, test_spectra.py)."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 30000))
    w = sps.get_window("hann", 203)
    got = np.concatenate(oracle.oaconvolve(x, w, 5000, -1, "same"), -1)
    assert np.allclose(got, sps.oaconvolve(x, w[None, :], mode="same", axes=-1))
    sos = sps.butter(4, [1, 100], btype="bandpass", fs=5000, output="sos")
    got = np.concatenate(oracle.sosfilt(x, sos, 7000, -1)[0], -1)
    assert np.allclose(got, sps.sosfilt(sos, x, axis=-1))
    cnt, f, p = oracle.welch_psd(x, 1000, -1, 0.5)
    rf, rp = sps.welch(x, fs=1000, nperseg=2000, noverlap=1000, window="hann", axis=-1)
    assert np.allclose(p, rp) and np.allclose(f, rf)


def test_protools_golden_bit_exact():
    """Masked producer + protools.mean / std / standardize."""
    g = golden("protools")
    x = signal(int(g["seed"]), int(g["rows"]), int(g["n"]), int(g["fs"])) + float(g["offset"])
    assert x.sum() == float(g["x_sum"])
    cs, mask = int(g["chunksize"]), g["mask"]
    for name, arr, axis in (("ax1", x, -1), ("ax0", np.ascontiguousarray(x.T), 0)):
        blocks = oracle.masked(arr, mask, cs, axis)
        assert [b.shape[axis] for b in blocks] == list(g["masked_lengths_" + name])
        assert np.array_equal(np.concatenate(blocks, axis), g["masked_" + name])
        chunks = oracle.split_chunks(arr, cs, axis)
        for ax in (0, 1):
            for keep in (0, 1):
                assert np.array_equal(oracle.pro_mean(chunks, axis, ax, keepdims=bool(keep)),
                                      g["mean_%s_axis%d_keep%d" % (name, ax, keep)])
                assert np.array_equal(oracle.pro_std(chunks, axis, ax, keepdims=bool(keep)),
                                      g["std_%s_axis%d_keep%d" % (name, ax, keep)])
            assert np.array_equal(np.concatenate(oracle.standardize(chunks, axis, ax), axis),
                                  g["standardized_%s_axis%d" % (name, ax)])
    xn = x.copy()
    for r, a, b in g["nan_spans"]:
        xn[r, a:b] = np.nan
    chunks = oracle.split_chunks(xn, cs, -1)
    with np.errstate(all="ignore"):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            assert np.array_equal(oracle.pro_mean(chunks, -1, -1), g["mean_nan"], equal_nan=True)
            assert np.array_equal(oracle.pro_std(chunks, -1, -1), g["std_nan"], equal_nan=True)


def test_hilbert_golden_bit_exact():
    g = golden("hilbert")
    x = _x(g)
    got = np.concatenate(oracle.oaconvolve(x, g["taps"], int(g["chunksize"]), -1, "same"), -1)
    assert np.array_equal(got, g["y"])
