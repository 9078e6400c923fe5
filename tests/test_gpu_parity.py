"""Parity of the sm_100a kernels, through the reference-facing operator API and
the C ABI, against the golden vectors of the real reference and the CPU oracle.
Run on the B200 box: `pytest -m gpu`."""

import numpy as np
import pytest
import scipy.signal as sps

import oracle
from openseize_b200 import producer

from tests import parity_cases as pc
from tests.conftest import has_cuda, relerr

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")]


@pytest.fixture(scope="module")
def dv():
    from openseize_b200.core import device

    device.require_cuda()
    return device


def _dev(dv, a):
    return dv.from_host(np.asarray(a, dtype=np.float64))


def test_native_library_is_what_runs(dv):
    from openseize_b200 import _abi

    before = _abi.launch_count()
    pc.fir_golden()
    assert _abi.launch_count() > before, "no kernel of libosz_b200.so was launched"


def test_fir_golden():
    pc.fir_golden()
    pc.fir_long_golden()
    pc.hilbert_golden()


def test_fir_oracle_sweep():
    pc.fir_oracle_sweep()


def test_narrow_input_dtypes():
    pc.narrow_input_dtypes()


@pytest.mark.parametrize("algo", [1, 2])
@pytest.mark.parametrize("ntaps", [1, 2, 15, 16, 31, 113, 400, 671, 1025])
def test_fir_kernels_vs_numpy(dv, algo, ntaps):
    """Both FIR kernels on ragged sizes: y = valid convolution of the halo'd span."""
    rng = np.random.default_rng(ntaps)
    taps = rng.standard_normal(ntaps)
    plan = dv.FirPlan(taps, algo)
    for rows, n_out in ((1, 1), (3, 1919), (2, 1921), (5, 10000), (2, 40001)):
        x = rng.standard_normal((rows, n_out + ntaps - 1))
        # odd leading dimension and an offset view: unaligned rows
        buf = dv.zeros((rows, x.shape[1] + 3))
        buf[:, 1:1 + x.shape[1]] = _dev(dv, x)
        y = plan.run(buf[:, 1:1 + x.shape[1]], n_out).cpu().numpy()
        ref = np.stack([np.convolve(r, taps, "valid") for r in x])
        assert relerr(y, ref) < 1e-12, (algo, ntaps, rows, n_out)


def test_fir_float32_compute(dv):
    """Opt-in float32 arithmetic of the overlap-save FIR (float64 in and out):
    within north_star's float32 tolerance, 1e-5 of the output peak, of the
    float64 oracle -- through the kernel and through the public operator."""
    import openseize_b200
    from openseize_b200.filtering.fir import Kaiser

    rng = np.random.default_rng(21)
    for fs in (5000, 30000):
        taps = Kaiser(500, 600, fs).coeffs
        plan = dv.FirPlan(taps, 3)
        assert plan.algo == 3
        for rows, n_out in ((1, 1), (3, 1919), (2, 40001)):
            x = rng.standard_normal((rows, n_out + len(taps) - 1)) + 0.5
            buf = dv.zeros((rows, x.shape[1] + 3))
            buf[:, 1:1 + x.shape[1]] = _dev(dv, x)
            y = plan.run(buf[:, 1:1 + x.shape[1]], n_out).cpu().numpy()
            ref = np.stack([np.convolve(r, taps, "valid") for r in x])
            assert relerr(y, ref) < 1e-5, (fs, rows, n_out)
    x = rng.standard_normal((4, 120000)) + 2.0
    filt = Kaiser(500, 600, 5000)
    ref = np.concatenate(oracle.oaconvolve(x, filt.coeffs, 30000, -1, "same"), -1)
    openseize_b200.set_compute("float32")
    try:
        y32 = filt(producer(x, 30000, -1), 30000, axis=-1).to_array()
    finally:
        openseize_b200.set_compute("float64")
    y64 = filt(producer(x, 30000, -1), 30000, axis=-1).to_array()
    assert y32.dtype == np.float64 and y32.shape == ref.shape
    assert 1e-9 < relerr(y32, ref) < 1e-5          # really float32 arithmetic, within tolerance
    assert relerr(y64, ref) < 1e-12


def test_fir_long_taps_block8192(dv):
    rng = np.random.default_rng(5)
    taps = rng.standard_normal(1500)
    plan = dv.FirPlan(taps, 2)
    x = rng.standard_normal((2, 30000 + 1499))
    y = plan.run(_dev(dv, x), 30000).cpu().numpy()
    ref = np.stack([np.convolve(r, taps, "valid") for r in x])
    assert relerr(y, ref) < 1e-12


def test_fir_partitioned_taps(dv):
    """More than 2049 taps: the taps are partitioned and the partial
    convolutions accumulated."""
    rng = np.random.default_rng(6)
    taps = rng.standard_normal(5000)
    plan = dv.FirPlan(taps)
    x = rng.standard_normal((2, 20000 + 4999))
    y = plan.run(_dev(dv, x), 20000).cpu().numpy()
    ref = np.stack([np.convolve(r, taps, "valid") for r in x])
    assert relerr(y, ref) < 1e-12


def test_iir_golden():
    pc.iir_golden()


def test_iir_oracle_sweep():
    pc.iir_oracle_sweep()


@pytest.mark.parametrize("per_thread", ["32", "16"])
@pytest.mark.parametrize("kind", ["butter8", "notch", "lowpass2"])
@pytest.mark.parametrize("n", [1, 16, 17, 32, 33, 4096, 4097, 8191, 8192, 8193, 50001])
def test_sos_kernel_vs_scipy(dv, n, kind, per_thread, monkeypatch):
    """Scan kernel vs the sequential DF2T recurrence: forward and reversed,
    random initial state, output and final state.  butter8 is C2's 8-section
    band-pass whose poles sit at radius 0.99978 (hard part 2).  Both
    kernel builds (32 and 16 samples per thread) are exercised."""
    monkeypatch.setenv("OSZ_SOS_T", per_thread)
    rng = np.random.default_rng(n)
    if kind == "butter8":
        sos = sps.butter(8, [1, 100], btype="bandpass", fs=5000, output="sos")
    elif kind == "notch":
        b, a = sps.iirnotch(60, 10, fs=30000)
        sos = np.concatenate([b, a])[None]
    else:
        sos = sps.butter(4, 300, fs=5000, output="sos")
    plan = dv.SosPlan(sos)
    rows = 3
    x = rng.standard_normal((rows, n)) + 1.0
    zi = rng.standard_normal((rows, sos.shape[0], 2))
    for reverse in (False, True):
        state = _dev(dv, zi)
        y = plan.run(_dev(dv, x), state, reverse=reverse).cpu().numpy()
        xr = x[:, ::-1] if reverse else x
        ry, rz = sps.sosfilt(sos, xr, axis=-1, zi=np.transpose(zi, (1, 0, 2)))
        ry = ry[:, ::-1] if reverse else ry
        assert relerr(y, ry) < 1e-10, (n, reverse)
        # the delay registers of these high-Q sections are ill-conditioned (|z| ~ 300 |y|)
        assert relerr(state.cpu().numpy(), np.transpose(rz, (1, 0, 2))) < 2e-8
        state2 = _dev(dv, zi)
        assert plan.run(_dev(dv, x), state2, reverse=reverse, want_output=False) is None
        # (a few rows on a big GPU: the full pass may run time-split, the state-only pass never)
        assert relerr(state2.cpu().numpy(), state.cpu().numpy()) < 2e-8


@pytest.mark.parametrize("kind", ["butter8", "notch", "butter2"])
@pytest.mark.parametrize("n", [131072, 300001, 1000000])
def test_sos_exact_time_split(dv, n, kind):
    """Few rows, long chunk: the row is cut into spans that run concurrently --
    two passes (span finals from rest -> combine with T^len -> true entering states)
    for long cascades, entering states as weighted sums of the `settle` samples before
    each span for one or two sections (notch: 3 / 12 spans at 300001 / 1e6 samples).
    The split must be invisible: same output and carried state as the sequential
    recurrence, forward and reversed."""
    rng = np.random.default_rng(n)
    if kind == "butter8":
        sos = sps.butter(8, [1, 100], btype="bandpass", fs=5000, output="sos")
    elif kind == "butter2":
        sos = sps.butter(2, [40, 60], btype="bandpass", fs=5000, output="sos")
        assert sos.shape[0] == 2
    else:
        b, a = sps.iirnotch(60, 10, fs=30000)
        sos = np.concatenate([b, a])[None]
    plan = dv.SosPlan(sos)          # 5 rows << SM count: the launch policy splits on its own
    rows = 5
    x = rng.standard_normal((rows, n)) + 1.0
    zi = rng.standard_normal((rows, sos.shape[0], 2))
    for reverse in (False, True):
        state = _dev(dv, zi)
        y = plan.run(_dev(dv, x), state, reverse=reverse).cpu().numpy()
        xr = x[:, ::-1] if reverse else x
        ry, rz = sps.sosfilt(sos, xr, axis=-1, zi=np.transpose(zi, (1, 0, 2)))
        ry = ry[:, ::-1] if reverse else ry
        assert relerr(y, ry) < 1e-10, (n, reverse)
        assert relerr(state.cpu().numpy(), np.transpose(rz, (1, 0, 2))) < 2e-8


@pytest.mark.parametrize("kernel", ["split", "scan", "seq", "auto"])
@pytest.mark.parametrize("design", ["butter3", "cheby5", "butter4bp", "ellip6", "butter7", "butter10"])
def test_tf_kernels_vs_scipy(dv, design, kernel, monkeypatch):
    """(b, a) filters above second order against scipy.signal.lfilter, on each kernel:
    spans running concurrently after a settle-length warm-up (sequential arithmetic), the
    scan over the companion matrix (orders 3..8; it re-associates the state through matrix
    powers, so it is held to 1e-7 here and only chosen by itself for benign designs), and
    one thread per row -- output, carried state, state-only pass, forward and reversed,
    lengths around the block sizes and long enough for many spans."""
    if kernel != "auto":
        monkeypatch.setenv("OSZ_TF_KERNEL", kernel)
    b, a = {"butter3": sps.butter(3, 0.2), "cheby5": sps.cheby1(5, 1, 0.25),
            "butter4bp": sps.butter(4, [0.05, 0.3], btype="bandpass"),
            "ellip6": sps.ellip(6, 0.5, 40, 0.3), "butter7": sps.butter(7, 0.15),
            "butter10": sps.butter(10, 0.4)}[design]
    k = max(len(a), len(b)) - 1
    plan = dv.TfPlan(b, a)
    rng = np.random.default_rng(k)
    tol = 1e-7 if kernel == "scan" else 1e-9
    for n in (1, 15, 16, 17, 4095, 4096, 4097, 9000, 50001, 400_003):
        x = rng.standard_normal((3, n)) + 0.5
        zi = rng.standard_normal((3, k)) * 0.1
        for reverse in (False, True):
            state = _dev(dv, zi)
            y = plan.run(_dev(dv, x), state, reverse=reverse).cpu().numpy()
            xr = x[:, ::-1] if reverse else x
            ry, rz = sps.lfilter(b, a, xr, axis=-1, zi=zi)
            ry = ry[:, ::-1] if reverse else ry
            assert relerr(y, ry) < tol, (design, n, reverse)
            assert relerr(state.cpu().numpy(), rz) < 100 * tol, (design, n, reverse)
            state2 = _dev(dv, zi)
            assert plan.run(_dev(dv, x), state2, reverse=reverse, want_output=False) is None
            assert relerr(state2.cpu().numpy(), rz) < 100 * tol


@pytest.mark.parametrize("tile", ["1", "0"])
@pytest.mark.parametrize("rows,n", [(1, 4096 * 70 + 5), (2, 4096 * 3), (32, 1_000_000), (300, 41_000),
                                    (7, 4097), (3, 2_000_003)])
def test_sos_tile_lookback(dv, rows, n, tile, monkeypatch):
    """One section: the tiled scan with a decoupled look-back (every 4096-sample tile
    scanned from rest, entering state composed from its predecessors' aggregates) against
    the sequential recurrence -- output, carried state, state-only pass, forward and
    reversed, float64 and float32 samples -- and the same cases on the one-CTA-per-row
    kernel it replaces (OSZ_SOS_TILE=0)."""
    monkeypatch.setenv("OSZ_SOS_TILE", tile)
    rng = np.random.default_rng(rows * 7 + n)
    b, a = sps.iirnotch(60, 10, fs=30000)
    sos = np.concatenate([b, a])[None]
    plan = dv.SosPlan(sos)
    x = rng.standard_normal((rows, n)) + 1.0
    zi = rng.standard_normal((rows, 1, 2))
    for reverse in (False, True):
        state = _dev(dv, zi)
        y = plan.run(_dev(dv, x), state, reverse=reverse).cpu().numpy()
        xr = x[:, ::-1] if reverse else x
        ry, rz = sps.sosfilt(sos, xr, axis=-1, zi=np.transpose(zi, (1, 0, 2)))
        ry = ry[:, ::-1] if reverse else ry
        assert relerr(y, ry) < 1e-10, (n, reverse)
        assert relerr(state.cpu().numpy(), np.transpose(rz, (1, 0, 2))) < 2e-8
        state2 = _dev(dv, zi)
        assert plan.run(_dev(dv, x), state2, reverse=reverse, want_output=False) is None
        assert relerr(state2.cpu().numpy(), np.transpose(rz, (1, 0, 2))) < 2e-8
    # the look-ahead entry point: start state zi * (first sample processed), state only
    zi0 = sps.sosfilt_zi(sos)
    for reverse in (True, False):
        got = plan.lookahead(zi0, _dev(dv, x), reverse=reverse).cpu().numpy()
        xr = x[:, ::-1] if reverse else x
        _, rz = sps.sosfilt(sos, xr, axis=-1, zi=zi0[:, None, :] * xr[None, :, :1])
        assert relerr(got, np.transpose(rz, (1, 0, 2))) < 2e-8
    import torch
    x32 = torch.from_numpy(x.astype(np.float32)).cuda()
    for reverse in (False, True):
        state = _dev(dv, zi)
        y32 = plan.run(x32, state, reverse=reverse).cpu().numpy()
        assert y32.dtype == np.float32
        xr = x.astype(np.float32).astype(np.float64)
        xr = xr[:, ::-1] if reverse else xr
        ry, rz = sps.sosfilt(sos, xr, axis=-1, zi=np.transpose(zi, (1, 0, 2)))
        ry = ry[:, ::-1] if reverse else ry
        assert np.max(np.abs(y32 - ry)) / np.max(np.abs(ry)) < 1e-6
        assert relerr(state.cpu().numpy(), np.transpose(rz, (1, 0, 2))) < 2e-8


def test_sos_tile_nan_inputs(dv):
    """NaNs in the samples -- including the all-ones NaN bit pattern, which is what the tile
    descriptors use for "not published" -- must propagate like in the sequential recurrence
    (everything after the first NaN of a row is NaN, rows without NaNs are untouched) and must
    not stall the look-back."""
    import torch

    b, a = sps.iirnotch(60, 10, fs=30000)
    sos = np.concatenate([b, a])[None]
    plan = dv.SosPlan(sos)
    rng = np.random.default_rng(11)
    rows, n = 6, 4096 * 12 + 576
    x = rng.standard_normal((rows, n))
    x[1, 30000] = np.nan
    x[3, 5] = np.inf
    x[4, 20000] = np.frombuffer(np.array([-1], dtype=np.int64).tobytes(), dtype=np.float64)[0]
    for reverse in (False, True):
        state = dv.zeros((rows, 1, 2))
        y = plan.run(torch.from_numpy(x).cuda(), state, reverse=reverse).cpu().numpy()
        xr = x[:, ::-1] if reverse else x
        with np.errstate(all="ignore"):
            ry, _ = sps.sosfilt(sos, xr, axis=-1, zi=np.zeros((1, rows, 2)))
        ry = ry[:, ::-1] if reverse else ry
        assert np.array_equal(np.isfinite(y), np.isfinite(ry)), reverse
        good = np.isfinite(ry)
        assert np.max(np.abs(y[good] - ry[good])) / np.max(np.abs(ry[good])) < 1e-10


@pytest.mark.parametrize("deal", ["0", "1"])
def test_sos_tile_random_shapes(dv, deal, monkeypatch):
    """Seeded sweep of the single-section scan over row counts, lengths, row pitches and
    element offsets (views into wider buffers: aligned ones take the TMA kernel, odd
    offsets / pitches the plain-load tile kernel or the row kernel), float64 and float32
    samples, forward and reversed, output written into a view as well -- both ways of
    dealing the tiles -- against the sequential recurrence."""
    import torch

    monkeypatch.setenv("OSZ_SOS_TILE_DEAL", deal)
    rng = np.random.default_rng(2024)
    b, a = sps.iirnotch(50, 8, fs=5000)
    sos = np.concatenate([b, a])[None]
    plan = dv.SosPlan(sos)
    for case in range(28):
        rows = int(rng.choice([1, 2, 5, 31, 64, 150, 300]))
        n = int(rng.choice([1, 17, 4095, 4096, 4097, 8192, 12289, 40_000, 100_003, 262_144]))
        if rows * n > 20_000_000:
            n = 40_000
        pad_l = int(rng.choice([0, 1, 2, 3, 16]))
        pitch = n + pad_l + int(rng.choice([0, 1, 5, 16]))
        f32 = bool(rng.integers(0, 2))
        reverse = bool(rng.integers(0, 2))
        npdt = np.float32 if f32 else np.float64
        xfull = (rng.standard_normal((rows, pitch)) + 0.3).astype(npdt)
        xdev = torch.from_numpy(xfull).cuda()
        ydev = torch.zeros_like(xdev)
        xv, yv = xdev[:, pad_l:pad_l + n], ydev[:, pad_l:pad_l + n]
        zi = rng.standard_normal((rows, 1, 2)) * 0.1
        state = _dev(dv, zi)
        plan.run(xv, state, reverse=reverse, out=yv)
        x64 = xfull[:, pad_l:pad_l + n].astype(np.float64)
        xr = x64[:, ::-1] if reverse else x64
        ry, rz = sps.sosfilt(sos, xr, axis=-1, zi=np.transpose(zi, (1, 0, 2)))
        ry = ry[:, ::-1] if reverse else ry
        got = ydev.cpu().numpy()
        tol = 2e-6 if f32 else 1e-10
        scale = max(np.max(np.abs(ry)), 1e-30)
        assert np.max(np.abs(got[:, pad_l:pad_l + n] - ry)) / scale < tol, (case, rows, n, pad_l, pitch, f32, reverse)
        # nothing outside the view was touched
        assert not got[:, :pad_l].any() and not got[:, pad_l + n:].any(), (case, rows, n)
        assert relerr(state.cpu().numpy(), np.transpose(rz, (1, 0, 2))) < 2e-7, (case, rows, n, f32)


@pytest.mark.parametrize("kind", ["notch", "butter2"])
def test_sos_tail_state(dv, kind):
    """State after a run of samples from rest as ONE weighted sum of the last `settle`
    samples (osz_sos_tail_state_f64) against the recurrence, forward and reversed."""
    rng = np.random.default_rng(17)
    if kind == "notch":
        b, a = sps.iirnotch(60, 10, fs=5000)
        sos = np.concatenate([b, a])[None]
    else:
        sos = sps.butter(2, [40, 60], btype="bandpass", fs=5000, output="sos")
    plan = dv.SosPlan(sos)
    assert plan.has_weights
    n = 3 * plan._settle + 1234
    x = rng.standard_normal((3, n)) + 0.5
    for reverse in (False, True):
        got = plan.tail_state(_dev(dv, x), reverse=reverse).cpu().numpy()
        xr = x[:, ::-1] if reverse else x
        _, zf = sps.sosfilt(sos, xr, axis=-1, zi=np.zeros((sos.shape[0], 3, 2)))
        assert relerr(got, np.transpose(zf, (1, 0, 2))) < 1e-10
    assert not dv.SosPlan(sps.butter(8, [1, 100], btype="bandpass", fs=5000,
                                     output="sos")).has_weights


def test_resample_golden():
    pc.resample_golden()


def test_resample_oracle_sweep():
    pc.resample_oracle_sweep()


@pytest.mark.parametrize("kernel", ["mma", "polyphase", "auto"])
@pytest.mark.parametrize("up,down,ntaps", [(1, 2, 41), (1, 20, 449), (1, 25, 561), (1, 60, 1301),
                                           (1, 300, 901), (3, 7, 155), (5, 1, 99), (2, 3, 64),
                                           (1, 25, 1231), (1, 3, 7), (1, 4, 200), (1, 16, 333),
                                           (1, 10, 1)])
def test_upfirdn_kernels_vs_scipy(dv, up, down, ntaps, kernel):
    """Decimating (tensor-core banded-Toeplitz and CUDA-core polyphase, R = 16 / 8 / 4
    tiles) and general kernels against one global resample_poly call, windows that
    start and end inside the recording."""
    rng = np.random.default_rng(ntaps)
    h = sps.firwin(ntaps, 1.0 / max(up, down))
    x = rng.standard_normal((3, 40000))
    ref = sps.resample_poly(x, up, down, axis=-1, window=h)
    try:
        plan = dv.UpfirdnPlan(h, up, down, kernel=kernel)
    except NotImplementedError:
        assert kernel == "mma"          # no tensor-core tile geometry for this filter
        pytest.skip("no tensor-core geometry for up=%d down=%d taps=%d" % (up, down, ntaps))
    if kernel == "mma":
        assert plan.kernel == "mma"
    if up > 1 or down == 1:
        assert plan.kernel == "general"
    elif kernel == "polyphase":
        assert plan.kernel in ("polyphase", "general")     # general: no tile fits (M = 60, 300)
    n_total = ref.shape[-1]
    y = plan.run(_dev(dv, x), 0, 0, n_total).cpu().numpy()
    assert relerr(y, ref) < 1e-12
    # a window of the recording: outputs whose support lies inside it
    lo, hi = 9000, 31000
    o_lo = ((lo + ntaps) * up) // down + 2
    o_hi = ((hi - ntaps) * up) // down - 2
    y = plan.run(_dev(dv, x[:, lo:hi]), lo, o_lo, o_hi - o_lo).cpu().numpy()
    assert relerr(y, ref[:, o_lo:o_hi]) < 1e-12


def test_float32_io_mode():
    pc.float32_io()


def test_analytic_golden():
    pc.analytic_golden()


def test_producer_tools_golden():
    pc.protools_golden()
    pc.masked_chain()
    pc.protools_edges()


def test_spectra_golden():
    pc.spectra_golden("pow2")
    pc.spectra_golden("nonpow2")          # nfft = 2000: generic mixed-radix path


def test_spectra_oracle_sweep():
    pc.spectra_oracle_sweep()


@pytest.mark.parametrize("nfft", [256, 512, 1024, 2048, 4096, 8192])
@pytest.mark.parametrize("detrend", ["constant", "linear", None])
def test_fft_sizes_vs_numpy(dv, nfft, detrend):
    """Every shared-memory FFT size through the STFT and Welch kernels."""
    rng = np.random.default_rng(nfft)
    rows, stride, nseg = 3, nfft // 2, 7
    x = rng.standard_normal((rows, (nseg - 1) * stride + nfft)) + 0.5
    w = sps.get_window("hann", nfft)
    norm = 1.0 / (1000.0 * np.sum(w ** 2))
    plan = dv.SpecPlan(nfft, stride, w, detrend, norm)
    X = plan.segments(_dev(dv, x), nseg, True).cpu().numpy()
    X = X[..., 0] + 1j * X[..., 1]
    psd_sum = dv.zeros((rows, nfft // 2 + 1))
    plan.welch_accum(_dev(dv, x), nseg, psd_sum)
    P = plan.segments(_dev(dv, x), nseg, False).cpu().numpy()
    ref_sum = 0
    for s in range(nseg):
        seg = x[:, s * stride:s * stride + nfft]
        if detrend:
            seg = sps.detrend(seg, axis=-1, type=detrend)
        R = np.fft.rfft(seg * w, axis=-1) * np.sqrt(norm)
        assert relerr(X[s], R) < 1e-12, (nfft, s)
        p = np.abs(R) ** 2
        p[:, 1:-1] *= 2
        assert relerr(P[s], p) < 1e-12
        ref_sum = ref_sum + p
    assert relerr(psd_sum.cpu().numpy(), ref_sum) < 1e-12


@pytest.mark.parametrize("nfft,stride", [(4096, 2048), (4096, 3687), (4096, 1024), (1024, 333),
                                         (512, 512), (2048, 1), (2400, 1200), (1000, 77),
                                         (10000, 5000)])
def test_welch_unaligned_and_strides(dv, nfft, stride):
    """Welch on rows that start at odd elements with an odd leading dimension
    (the TMA span fetch copies from the aligned address below and patches odd
    tails), strides other than nfft/2, and 1 .. 41 segments: the ping-pong
    kernel (powers of two), the mixed-radix kernel (2400, 1000, 10000)."""
    rng = np.random.default_rng(nfft + stride)
    w = sps.get_window("hamming", nfft)
    norm = 1.0 / (1000.0 * np.sum(w ** 2))
    plan = dv.SpecPlan(nfft, stride, w, "constant", norm)
    for rows, nseg in ((1, 1), (5, 2), (2, 3), (3, 8), (2, 41)):
        if stride == 1 and nseg > 8:
            continue
        n = (nseg - 1) * stride + nfft
        x = rng.standard_normal((rows, n)) + 2.0
        buf = dv.zeros((rows, n + 5))
        buf[:, 3:3 + n] = _dev(dv, x)
        psd_sum = dv.zeros((rows, nfft // 2 + 1))
        plan.welch_accum(buf[:, 3:3 + n], nseg, psd_sum)
        ref = 0
        for k in range(nseg):
            seg = sps.detrend(x[:, k * stride:k * stride + nfft], axis=-1, type="constant")
            p = np.abs(np.fft.rfft(seg * w, axis=-1)) ** 2 * norm
            p[:, 1:-1] *= 2
            ref = ref + p
        assert relerr(psd_sum.cpu().numpy(), ref) < 1e-11, (nfft, stride, rows, nseg)


@pytest.mark.parametrize("nfft,stride", [(4096, 2048), (4096, 3687), (2048, 1024), (1024, 333),
                                         (512, 512), (1024, 512), (512, 256)])
@pytest.mark.parametrize("detrend", ["constant", "linear", None])
def test_welch_float32_compute(dv, nfft, stride, detrend):
    """Opt-in float32 arithmetic of the Welch accumulation (float64 samples in,
    float64 sums out): within north_star's float32 tolerance, 1e-5 of the
    largest bin, of the float64 result -- with a DC offset 1000x the signal
    (samples are centred in float64 before they are narrowed), a drift,
    unaligned rows and 1 .. 41 segments."""
    rng = np.random.default_rng(nfft + stride)
    w = sps.get_window("hann", nfft)
    norm = 1.0 / (1000.0 * np.sum(w ** 2))
    plan = dv.SpecPlan(nfft, stride, w, detrend, norm, "float32")
    assert plan.compute == "float32"
    for rows, nseg in ((1, 1), (5, 2), (2, 3), (3, 8), (2, 41)):
        n = (nseg - 1) * stride + nfft
        x = rng.standard_normal((rows, n)) + 3 * np.sin(2 * np.pi * 0.01 * np.arange(n))
        if detrend:
            x = x + 1000.0 + (np.arange(n) * (2.0 / nfft) if detrend == "linear" else 0.0)
        buf = dv.zeros((rows, n + 5))
        buf[:, 3:3 + n] = _dev(dv, x)
        psd_sum = dv.zeros((rows, nfft // 2 + 1))
        plan.welch_accum(buf[:, 3:3 + n], nseg, psd_sum)
        ref = 0
        for k in range(nseg):
            seg = x[:, k * stride:k * stride + nfft]
            if detrend:
                seg = sps.detrend(seg, axis=-1, type=detrend)
            p = np.abs(np.fft.rfft(seg * w, axis=-1)) ** 2 * norm
            p[:, 1:-1] *= 2
            ref = ref + p
        got = psd_sum.cpu().numpy()
        err = np.max(np.abs(got - ref), axis=-1) / np.max(ref, axis=-1)    # per channel
        assert np.all(err < 1e-5), (nfft, stride, detrend, rows, nseg, err)


def test_psd_float32_compute(dv):
    """psd() under set_compute("float32"): really float32 arithmetic, within tolerance."""
    import openseize_b200
    from openseize_b200.spectra.estimators import psd

    rng = np.random.default_rng(77)
    fs = 4096.0
    t = np.arange(400000) / fs
    x = rng.standard_normal((3, t.size)) + 20 * np.sin(2 * np.pi * 60 * t) + 500.0
    cnt_o, f_o, ref = oracle.welch_psd(x, fs, -1, fs / 4096)
    openseize_b200.set_compute("float32")
    try:
        cnt32, f32, p32 = psd(producer(x, 100000, -1), fs, resolution=fs / 4096)
    finally:
        openseize_b200.set_compute("float64")
    cnt64, f64, p64 = psd(producer(x, 100000, -1), fs, resolution=fs / 4096)
    assert cnt32 == cnt64 == cnt_o and np.array_equal(f32, f_o)
    e32 = np.max(np.abs(p32 - ref), axis=-1) / np.max(ref, axis=-1)
    e64 = np.max(np.abs(p64 - ref), axis=-1) / np.max(ref, axis=-1)
    assert np.all(e64 < 1e-11)
    assert np.all(e32 < 1e-5) and np.any(e32 > 1e-11)
    # plans outside the float32 kernel's range keep float64 and say so
    w = sps.get_window("hann", 2000)
    assert dv.SpecPlan(2000, 1000, w, "constant", 1.0, "float32").compute == "float64"


@pytest.mark.parametrize("nfft", [2, 3, 60, 97, 250, 1001, 1009, 2000, 10000, 16384])
def test_generic_nfft_vs_numpy(dv, nfft):
    """Mixed-radix (2..16, 3, 5, 7, 11, 13) and Bluestein (97, 1009) lengths,
    odd and even, through the periodogram / STFT / Welch entry points."""
    rng = np.random.default_rng(nfft)
    rows, nseg = 3, 5
    stride = max(1, nfft // 2)
    x = rng.standard_normal((rows, (nseg - 1) * stride + nfft)) + 0.5
    w = sps.get_window("hann", nfft) if nfft > 3 else np.ones(nfft)
    norm = 1.0 / (1000.0 * np.sum(w ** 2))
    for detrend in ("constant", "linear", None):
        if detrend == "linear" and nfft < 3:
            continue
        plan = dv.SpecPlan(nfft, stride, w, detrend, norm)
        assert plan.path == 2
        X = plan.segments(_dev(dv, x), nseg, True).cpu().numpy()
        X = X[..., 0] + 1j * X[..., 1]
        P = plan.segments(_dev(dv, x), nseg, False).cpu().numpy()
        psd_sum = dv.zeros((rows, nfft // 2 + 1))
        plan.welch_accum(_dev(dv, x), nseg, psd_sum)
        ref_sum = 0
        for s in range(nseg):
            seg = x[:, s * stride:s * stride + nfft]
            if detrend:
                seg = sps.detrend(seg, axis=-1, type=detrend)
            R = np.fft.rfft(seg * w, axis=-1) * np.sqrt(norm)
            scale = max(np.max(np.abs(R)), 1e-300)
            assert np.max(np.abs(X[s] - R)) / scale < 1e-11, (nfft, detrend, s)
            p = np.abs(R) ** 2
            if nfft % 2:
                p[:, 1:] *= 2
            else:
                p[:, 1:-1] *= 2
            assert np.max(np.abs(P[s] - p)) / max(np.max(p), 1e-300) < 1e-11
            ref_sum = ref_sum + p
        got = psd_sum.cpu().numpy()
        assert np.max(np.abs(got - ref_sum)) / max(np.max(ref_sum), 1e-300) < 1e-11


def test_default_resolution_psd():
    """psd() with the reference's default resolution (0.5 Hz -> nfft = 2 fs)."""
    pc.spectra_oracle_sweep(fs=1000, resolutions=(0.5,))


def test_pipeline_chain_on_device():
    pc.pipeline_chain()


def test_fused_fir_decimate():
    pc.fused_fir_decimate()


def test_fused_iir_fir_decimate():
    pc.fused_iir_fir_decimate()


def test_c5_real_parameters():
    """The bench's own configuration at its real parameters (30 kHz notch with the
    shortened look-ahead active, 671-tap FIR, fused 1231-tap M=25 decimator, nfft
    4096 on the 1200 Hz stream, chunksize 1e6, 6 chunks + ragged tail)."""
    errs = pc.c5_real_parameters()
    assert max(errs.values()) <= 1e-9, errs


def test_full_size_properties(dv):
    """BASELINE-sized chunk (64 rows x 1e6): size-independent properties --
    impulse response reproduces the taps, linearity, filter-then-PSD of a sine
    peaks at the sine's bin, forward scan == two half-length scans."""
    rng = np.random.default_rng(3)
    rows, n = 64, 1_000_000
    from openseize_b200.filtering.fir import Kaiser

    taps = Kaiser(500, 600, 5000).coeffs
    k = len(taps)
    plan = dv.FirPlan(taps)
    x = dv.zeros((rows, n + k - 1))
    x[:, k - 1 + 1000] = 1.0
    y = plan.run(x, n)
    got = y[:, 1000:1000 + k].cpu().numpy()
    assert relerr(got, np.broadcast_to(taps, got.shape)) < 1e-13
    assert float(y[:, :1000].abs().max()) < 1e-15 and float(y[:, 1000 + k:].abs().max()) < 1e-15
    a = _dev(dv, rng.standard_normal((rows, 200000 + k - 1)))
    b = _dev(dv, rng.standard_normal((rows, 200000 + k - 1)))
    lin = plan.run(2.0 * a - 3.0 * b, 200000) - (2.0 * plan.run(a, 200000) - 3.0 * plan.run(b, 200000))
    assert float(lin.abs().max()) < 1e-12
    sos = sps.butter(8, [1, 100], btype="bandpass", fs=5000, output="sos")
    sp = dv.SosPlan(sos)
    xs = _dev(dv, rng.standard_normal((rows, n)))
    st = dv.zeros((rows, 8, 2))
    whole = sp.run(xs, st)
    st2 = dv.zeros((rows, 8, 2))
    h1 = sp.run(xs[:, :n // 2 + 7], st2)
    h2 = sp.run(xs[:, n // 2 + 7:], st2)
    two = dv.cat_time([h1, h2])
    assert float((whole - two).abs().max()) / float(whole.abs().max()) < 1e-10
    assert relerr(st2.cpu().numpy(), st.cpu().numpy()) < 1e-9


def test_baseline_configs_full_length():
    """BASELINE.json configs 1-4 at their FULL recording length against the
    oracle, on as many channels as the oracle finishes in seconds (channels are
    independent): C1 complete (4 x 18 000 000, the reference's own
    CPU-runnable case), C2 on 2 of 64 channels, C3 and C4 on one channel."""
    from openseize_b200.filtering.fir import Kaiser
    from openseize_b200.filtering.iir import Butter
    from openseize_b200.resampling.resampling import downsample
    from openseize_b200.spectra.estimators import psd

    cs = 1_000_000
    rng = np.random.default_rng([1, 0])
    # C1: Kaiser 500/600 @ 5 kHz, 113 taps, mode 'same', chunksize 1e6
    x = rng.standard_normal((4, 18_000_000))
    filt = Kaiser(fpass=500, fstop=600, fs=5000)
    y = filt(producer(x, cs, -1), cs, axis=-1, mode="same").to_array()
    ref = np.concatenate(oracle.oaconvolve(x, filt.coeffs, cs, -1, "same"), -1)
    assert y.shape == ref.shape == x.shape and relerr(y, ref) < 1e-9
    # C2: Butterworth band-pass 1-100 Hz (8 sections), forward-backward, 18 chunks
    butter = Butter(fpass=[1, 100], fstop=[0.5, 200], fs=5000, gpass=1, gstop=40)
    z = butter(producer(x[:2], cs, -1), cs, axis=-1, dephase=True).to_array()
    ref = np.concatenate(oracle.sosfiltfilt(x[:2], butter.coeffs, cs, -1), -1)
    assert z.shape == ref.shape and relerr(z, ref) < 1e-9
    del y, z, ref
    # C3: downsample 5000 -> 250 Hz, 108 000 000 samples: lengths exact, values 1e-9
    x1 = np.random.default_rng([1, 1]).standard_normal((1, 108_000_000))
    d = downsample(producer(x1, cs, -1), M=20, fs=5000, chunksize=cs, axis=-1)
    got = list(d)
    refl = oracle.polyphase_resample(x1, 1, 20, 5000, cs, -1)
    assert d.shape == (1, 5_400_000) and sum(a.shape[-1] for a in got) == 5_400_000
    assert relerr(np.concatenate(got, -1), np.concatenate(refl, -1)) < 1e-9
    del got, refl, d
    # C4: Welch PSD nfft 4096, 50 % overlap, Hann @ 30 kHz: 52 733 segments
    cnt, freqs, est = psd(producer(x1, cs, -1), fs=30000, resolution=30000 / 4096)
    ocnt, ofreqs, oest = oracle.welch_psd(x1, 30000, -1, 30000 / 4096)
    assert cnt == ocnt == 52733 and np.array_equal(freqs, ofreqs)
    assert relerr(est, oest) < 1e-9


def test_decimator_float32_compute(dv):
    """Opt-in float32 arithmetic of the decimating polyphase filter (float64 in
    and out): kernel level against scipy, and through downsample() -- alone and
    fused with an upstream FIR -- within 1e-5 of the output peak."""
    import openseize_b200
    from openseize_b200.filtering.fir import Kaiser
    from openseize_b200.resampling.resampling import downsample

    rng = np.random.default_rng(31)
    for fs, M in ((5000, 20), (30000, 25), (1000, 2), (1000, 7)):
        h = oracle.resample_filter(1, M, fs)
        plan = dv.UpfirdnPlan(h, 1, M, "float32")
        assert plan.compute == "float32"
        for rows, n in ((1, 5 * M + 3), (3, 40000 + M - 1)):
            x = rng.standard_normal((rows, n)) + 0.5
            ref = sps.resample_poly(x, 1, M, axis=-1, window=h)
            y = plan.run(_dev(dv, x), 0, 0, ref.shape[1]).cpu().numpy()
            assert relerr(y, ref) < 1e-5, (fs, M, rows, n)
    assert dv.UpfirdnPlan(oracle.resample_filter(3, 7, 1000), 3, 7, "float32").compute == "float64"
    fs, cs = 5000, 30000
    x = rng.standard_normal((4, 150000)) + 2.0
    filt = Kaiser(500, 600, fs)
    r1 = np.concatenate(oracle.oaconvolve(x, filt.coeffs, cs, -1, "same"), -1)
    ref_plain = np.concatenate(oracle.polyphase_resample(x, 1, 4, fs, cs, -1), -1)
    ref_fused = np.concatenate(oracle.polyphase_resample(r1, 1, 4, fs, cs, -1), -1)
    openseize_b200.set_compute("float32")
    try:
        d32 = downsample(producer(x, cs, -1), 4, fs, cs, axis=-1).to_array()
        f32 = downsample(filt(producer(x, cs, -1), cs, axis=-1), 4, fs, cs, axis=-1).to_array()
    finally:
        openseize_b200.set_compute("float64")
    d64 = downsample(producer(x, cs, -1), 4, fs, cs, axis=-1).to_array()
    assert d32.shape == ref_plain.shape and f32.shape == ref_fused.shape
    assert 1e-10 < relerr(d32, ref_plain) < 1e-5 and relerr(f32, ref_fused) < 1e-5
    assert relerr(d64, ref_plain) < 1e-12


@pytest.mark.parametrize("nfft", [512, 1024, 2048, 4096])
@pytest.mark.parametrize("detrend", ["constant", "linear", None])
def test_segments_float32_compute(dv, nfft, detrend):
    """Opt-in float32 arithmetic of the per-segment kernels (STFT / periodogram):
    float64 samples in, complex128 / float64 out, within 1e-5 of the peak."""
    rng = np.random.default_rng(nfft + 5)
    rows, stride, nseg = 3, nfft // 2, 7
    x = rng.standard_normal((rows, (nseg - 1) * stride + nfft + 2)) + (300.0 if detrend else 0.5)
    w = sps.get_window("hann", nfft)
    norm = 1.0 / (1000.0 * np.sum(w ** 2))
    plan = dv.SpecPlan(nfft, stride, w, detrend, norm, "float32")
    assert plan.compute == "float32"
    xd = _dev(dv, x)[:, 1:]                                  # rows start at an odd element
    X = plan.segments(xd, nseg, True).cpu().numpy()
    X = X[..., 0] + 1j * X[..., 1]
    P = plan.segments(xd, nseg, False).cpu().numpy()
    for s in range(nseg):
        seg = x[:, 1 + s * stride:1 + s * stride + nfft]
        if detrend:
            seg = sps.detrend(seg, axis=-1, type=detrend)
        R = np.fft.rfft(seg * w, axis=-1) * np.sqrt(norm)
        p = np.abs(R) ** 2
        p[:, 1:-1] *= 2
        assert 1e-10 < relerr(X[s], R) < 1e-5, (nfft, detrend, s)
        assert relerr(P[s], p) < 1e-5


def test_stft_float32_compute(dv):
    import openseize_b200
    from openseize_b200.spectra.estimators import stft

    rng = np.random.default_rng(78)
    fs = 2048.0
    x = rng.standard_normal((2, 50000)) + 40.0
    of, ot, oX = oracle.stft(x, fs, -1, fs / 1024)
    openseize_b200.set_compute("float32")
    try:
        f, t, X = stft(producer(x, 20000, -1), fs, resolution=fs / 1024)
    finally:
        openseize_b200.set_compute("float64")
    assert np.array_equal(f, of) and np.array_equal(t, ot) and X.shape == oX.shape
    assert X.dtype == np.complex128 and 1e-10 < relerr(X, oX) < 1e-5
