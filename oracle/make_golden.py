"""Generate tests/golden/*.npz from the REAL reference and pin the oracle.

TEST INFRASTRUCTURE ONLY.  Run in the build container, where the read-only
reference checkout exists:

    python oracle/make_golden.py            # writes tests/golden/*.npz

For every hot-path operator this script
  1. builds a small seeded input,
  2. runs the unmodified reference (imported from /root/reference/src with
     matplotlib stubbed) through its public operator API,
  3. runs the oracle restatement on the same input and asserts agreement
     (bit-exact where the oracle calls the same compiled routine in the same
     order, <= 1e-12 of peak otherwise),
  4. stores the reference's outputs (and the recipe of the input) as a golden
     vector.

The GPU box has no /root/reference; tests there read only the .npz files.
"""

import os
import sys
import types

import numpy as np

REF_SRC = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")


def _import_reference():
    """Put the reference on sys.path with the plotting imports stubbed."""
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches",
                 "matplotlib.widgets"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib.pyplot"].Axes = object
    sys.modules["matplotlib.patches"].Rectangle = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].patches = sys.modules["matplotlib.patches"]
    sys.path.insert(0, REF_SRC)
    import openseize  # noqa: F401
    from openseize import producer
    from openseize.core import numerical as nm
    from openseize.filtering import fir, iir
    from openseize.resampling import resampling
    from openseize.spectra import estimators
    return producer, nm, fir, iir, resampling, estimators


def signal(seed, rows, n, fs):
    """Seeded test signal with line noise, an alpha rhythm and drift
    (SURVEY.md section 8d parity variant)."""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / fs
    x = rng.standard_normal((rows, n))
    x += 20 * np.sin(2 * np.pi * 60 * t) + 30 * np.sin(2 * np.pi * 8 * t) + 5 * t / t[-1]
    return x


def relerr(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def protools_golden(report):
    """Masked producers (core/producer.py:379-444) and protools.mean / std /
    standardize (core/protools.py:500-668): SURVEY 8f, N3."""
    import oracle
    from openseize import producer
    from openseize.core import protools

    fs = 500
    x = signal(51, 3, 4500, fs) + 40.0
    rng = np.random.default_rng(52)
    # a "sleep state" mask: runs of kept / dropped samples, some chunks fully dropped
    mask = np.repeat(rng.random(45) < 0.55, 100)
    mask[1000:2100] = False
    gold = {"seed": 51, "rows": 3, "n": 4500, "fs": fs, "offset": 40.0, "x_sum": x.sum(),
            "mask": mask, "chunksize": 500}
    for name, arr, axis in (("ax1", x, -1), ("ax0", np.ascontiguousarray(x.T), 0)):
        mpro = producer(arr, 500, axis, mask=mask)
        ref_list = [np.array(a) for a in mpro]
        mine_list = oracle.masked(arr, mask, 500, axis)
        assert [a.shape for a in ref_list] == [a.shape for a in mine_list]
        assert all(np.array_equal(a, b) for a, b in zip(ref_list, mine_list))
        gold["masked_lengths_" + name] = np.array([a.shape[axis] for a in ref_list])
        gold["masked_" + name] = np.concatenate(ref_list, axis)
        chunks = oracle.split_chunks(arr, 500, axis)
        for ax in (0, 1):
            for keep in (False, True):
                m = protools.mean(producer(arr, 500, axis), ax, keepdims=keep)
                s = protools.std(producer(arr, 500, axis), ax, keepdims=keep)
                om = oracle.pro_mean(chunks, axis, ax, keepdims=keep)
                osd = oracle.pro_std(chunks, axis, ax, keepdims=keep)
                assert np.array_equal(m, om) and np.array_equal(s, osd), (name, ax, keep)
                gold["mean_%s_axis%d_keep%d" % (name, ax, keep)] = m
                gold["std_%s_axis%d_keep%d" % (name, ax, keep)] = s
            z = protools.standardize(producer(arr, 500, axis), ax).to_array()
            oz = np.concatenate(oracle.standardize(chunks, axis, ax), axis)
            assert np.array_equal(z, oz), (name, ax)
            gold["standardized_%s_axis%d" % (name, ax)] = z
    # NaNs: the chunk-weighted nanmean of the reference (n * nanmean per chunk)
    xn = x.copy()
    xn[0, 100:160] = np.nan
    xn[2, 2500:3400] = np.nan
    chunks = oracle.split_chunks(xn, 500, -1)
    m = protools.mean(producer(xn, 500, -1), -1)
    s = protools.std(producer(xn, 500, -1), -1)
    assert np.array_equal(m, oracle.pro_mean(chunks, -1, -1), equal_nan=True)
    assert np.array_equal(s, oracle.pro_std(chunks, -1, -1), equal_nan=True)
    gold["nan_spans"] = np.array([[0, 100, 160], [2, 2500, 3400]])
    gold["mean_nan"], gold["std_nan"] = m, s
    np.savez_compressed(os.path.join(GOLD, "protools.npz"), **gold)
    report.append(("protools", "bit-exact vs reference (masked producer, mean, std, standardize)"))


def hilbert_golden(report):
    """Type III FIR Hilbert transformer (filtering/special.py:16-133)."""
    import oracle
    from openseize import producer
    from openseize.filtering.special import Hilbert

    fs = 500
    x = signal(71, 2, 6000, fs)
    hil = Hilbert(width=12, fs=fs)
    ref = hil(producer(x, 1500, -1), 1500, axis=-1).to_array()
    mine = np.concatenate(oracle.oaconvolve(x, hil.coeffs, 1500, -1, "same"), -1)
    assert np.array_equal(ref, mine)
    np.savez_compressed(os.path.join(GOLD, "hilbert.npz"), seed=71, rows=2, n=6000, fs=fs,
                        width=12, taps=hil.coeffs, chunksize=1500, x_sum=x.sum(), y=ref)
    report.append(("hilbert", "bit-exact vs reference"))


def analytic_golden(report):
    """Analytic signal, amplitudes and phases (experimental/coupling/transforms.py:
    107-192)."""
    from openseize.experimental.coupling.transforms import Analytic

    fs = 400
    x = signal(81, 2, 9000, fs)
    ana = Analytic(x, fs, chunksize=2500, axis=-1, width=4)
    z = ana.signal.to_array()
    amp = ana.amplitudes.to_array()
    ph = ana.phases.to_array()
    assert z.dtype == np.complex128 and z.shape == x.shape
    np.savez_compressed(os.path.join(GOLD, "analytic.npz"), seed=81, rows=2, n=9000, fs=fs,
                        width=4, chunksize=2500, x_sum=x.sum(), z=z, amplitudes=amp, phases=ph)
    report.append(("analytic", "reference outputs stored"))


def main():
    sys.path.insert(0, ROOT)
    import oracle
    producer, nm, fir, iir, resampling, estimators = _import_reference()
    os.makedirs(GOLD, exist_ok=True)
    report = []
    only = {"protools": protools_golden, "hilbert": hilbert_golden, "analytic": analytic_golden}
    if sys.argv[1:] and all(a in only for a in sys.argv[1:]):   # only the named fixtures
        for a in sys.argv[1:]:
            only[a](report)
        print(report)
        return

    # ---------------- FIR: Kaiser 113 taps, all modes, both axes -----------
    fs = 5000
    kais = fir.Kaiser(fpass=500, fstop=600, fs=fs)
    x = signal(11, 3, 9000, fs)
    gold = {"seed": 11, "rows": 3, "n": 9000, "fs": fs, "taps": kais.coeffs,
            "chunksize": 2500, "x_sum": x.sum()}
    for mode in ("same", "full", "valid"):
        ref = kais(producer(x, 2500, -1), 2500, axis=-1, mode=mode).to_array()
        mine = np.concatenate(oracle.oaconvolve(x, kais.coeffs, 2500, -1, mode), -1)
        assert ref.shape == mine.shape, (mode, ref.shape, mine.shape)
        assert np.array_equal(ref, mine), (mode, relerr(mine, ref))
        # and the mathematical definition (numpy convolve per row)
        direct = np.stack([np.convolve(r, kais.coeffs, mode) for r in x])
        assert relerr(ref, direct) < 1e-12
        gold["y_" + mode] = ref
    xt = np.ascontiguousarray(x[:2].T)                       # axis=0 layout
    ref0 = kais(producer(xt, 2500, 0), 2500, axis=0, mode="same").to_array()
    mine0 = np.concatenate(oracle.oaconvolve(xt, kais.coeffs, 2500, 0, "same"), 0)
    assert np.array_equal(ref0, mine0)
    gold["y_same_axis0"] = ref0
    # raw generator block structure (yield lengths)
    raw = [a.shape[-1] for a in nm.oaconvolve(producer(x, 2500, -1), kais.coeffs, -1, "same")]
    mine_raw = [a.shape[-1] for a in oracle.oaconvolve(x, kais.coeffs, 2500, -1, "same")]
    assert raw == mine_raw, (raw, mine_raw)
    gold["raw_block_lengths_same"] = np.array(raw)
    np.savez_compressed(os.path.join(GOLD, "fir_kaiser113.npz"), **gold)
    report.append(("fir_kaiser113", "bit-exact vs reference; <1e-12 vs np.convolve"))

    # long filter: Kaiser(500, 600, fs=30000) -> 671 taps
    fs = 30000
    k671 = fir.Kaiser(fpass=500, fstop=600, fs=fs)
    assert len(k671.coeffs) == 671
    x = signal(12, 2, 20000, fs)
    ref = k671(producer(x, 6000, -1), 6000, axis=-1, mode="same").to_array()
    mine = np.concatenate(oracle.oaconvolve(x, k671.coeffs, 6000, -1, "same"), -1)
    assert np.array_equal(ref, mine)
    np.savez_compressed(os.path.join(GOLD, "fir_kaiser671.npz"), seed=12, rows=2,
                        n=20000, fs=fs, taps=k671.coeffs, chunksize=6000,
                        x_sum=x.sum(), y_same=ref)
    report.append(("fir_kaiser671", "bit-exact vs reference"))

    # ---------------- IIR: Butterworth bandpass SOS (8 sections) -----------
    fs = 5000
    butter = iir.Butter(fpass=[1, 100], fstop=[0.5, 200], fs=fs, gpass=1, gstop=40)
    sos = butter.coeffs
    assert sos.shape == (8, 6), sos.shape
    x = signal(21, 2, 12000, fs)
    gold = {"seed": 21, "rows": 2, "n": 12000, "fs": fs, "sos": sos, "x_sum": x.sum()}
    for cs in (4000, 5000):
        ref = butter(producer(x, cs, -1), cs, axis=-1, dephase=True).to_array()
        mine = np.concatenate(oracle.sosfiltfilt(x, sos, cs, -1), -1)
        assert np.array_equal(ref, mine), relerr(mine, ref)
        gold["y_filtfilt_cs%d" % cs] = ref
    ref = butter(producer(x, 4000, -1), 4000, axis=-1, dephase=False).to_array()
    mine = np.concatenate(oracle.sosfilt(x, sos, 4000, -1)[0], -1)
    assert np.array_equal(ref, mine)
    gold["y_fwd_cs4000"] = ref
    np.savez_compressed(os.path.join(GOLD, "iir_butter8.npz"), **gold)
    report.append(("iir_butter8", "bit-exact vs reference"))

    # ---------------- IIR: Notch (b, a) -----------------------------------
    fs = 5000
    notch = iir.Notch(fstop=60, width=6, fs=fs)
    b, a = notch.coeffs
    x = signal(22, 2, 12000, fs)
    ref = notch(producer(x, 5000, -1), 5000, axis=-1, dephase=True).to_array()
    mine = np.concatenate(oracle.filtfilt(x, (b, a), 5000, -1), -1)
    assert np.array_equal(ref, mine)
    reff = notch(producer(x, 5000, -1), 5000, axis=-1, dephase=False).to_array()
    minef = np.concatenate(oracle.lfilter(x, (b, a), 5000, -1)[0], -1)
    assert np.array_equal(reff, minef)
    np.savez_compressed(os.path.join(GOLD, "iir_notch60.npz"), seed=22, rows=2,
                        n=12000, fs=fs, b=b, a=a, chunksize=5000, x_sum=x.sum(),
                        y_filtfilt=ref, y_fwd=reff)
    report.append(("iir_notch60", "bit-exact vs reference"))

    # ---------------- resampling -----------------------------------------
    fs = 5000
    x = signal(31, 2, 30000, fs)
    gold = {"seed": 31, "rows": 2, "n": 30000, "fs": fs, "x_sum": x.sum()}
    xfull = x
    for name, (L, M, cs, nuse) in {"down20": (1, 20, 7000, 30000),
                                   "up2": (2, 1, 2500, 9000),
                                   "rs3_7": (3, 7, 4100, 30000)}.items():
        x = xfull[:, :nuse]
        ref_pro = resampling.resample(producer(x, cs, -1), L, M, fs, cs, axis=-1)
        ref_list = [np.array(arr) for arr in ref_pro]
        ref = np.concatenate(ref_list, -1)
        raw = [arr.shape[-1] for arr in
               nm.polyphase_resample(producer(x, cs, -1), L, M, fs, fir.Kaiser, -1)]
        mine_list = oracle.polyphase_resample(x, L, M, fs, cs, -1)
        assert [m.shape[-1] for m in mine_list] == raw, (name, raw)
        mine = np.concatenate(mine_list, -1)
        assert ref.shape == mine.shape == ref_pro.shape, (name, ref.shape, mine.shape)
        assert np.array_equal(ref, mine), (name, relerr(mine, ref))
        h = oracle.resample_filter(L, M, fs)
        import scipy.signal as sps
        glob = sps.resample_poly(x, L, M, axis=-1, window=h)
        assert np.array_equal(glob, ref), name
        gold["y_" + name] = ref
        gold["raw_" + name] = np.array(raw)
        gold["h_" + name] = h
        gold["LMcsn_" + name] = np.array([L, M, cs, nuse])
    np.savez_compressed(os.path.join(GOLD, "resample.npz"), **gold)
    report.append(("resample", "bit-exact vs reference and vs global resample_poly"))

    # ---------------- Welch PSD / STFT -------------------------------------
    for name, fs, res, n in (("pow2", 1024, 1.0, 30000), ("nonpow2", 1000, 0.5, 30000)):
        x = signal(41, 2, n, fs)
        gold = {"seed": 41, "rows": 2, "n": n, "fs": fs, "resolution": res,
                "x_sum": x.sum()}
        for det in ("constant", "linear"):
            for scal in ("density", "spectrum"):
                cnt, f, p = estimators.psd(producer(x, 5000, -1), fs, axis=-1,
                                           resolution=res, detrend=det, scaling=scal)
                mcnt, mf, mp = oracle.welch_psd(x, fs, -1, res, detrend=det, scaling=scal)
                assert cnt == mcnt and np.array_equal(f, mf)
                assert np.array_equal(p, mp), relerr(mp, p)
                gold["psd_%s_%s" % (det, scal)] = p
                gold["psd_cnt"] = cnt
                gold["freqs"] = f
        for bnd in (True, False):
            for pad in (True, False):
                f, t, X = estimators.stft(producer(x, 5000, -1), fs, axis=-1,
                                          resolution=res, boundary=bnd, padded=pad)
                mf, mt, mX = oracle.stft(x, fs, -1, res, boundary=bnd, padded=pad)
                assert np.array_equal(f, mf) and np.array_equal(t, mt)
                assert X.shape == mX.shape and np.array_equal(X, mX)
                key = "stft_b%d_p%d" % (bnd, pad)
                gold[key + "_time"] = t
                # keep fixtures small: first 3 and last 2 segments only
                idx = np.r_[0:3, X.shape[-1] - 2:X.shape[-1]]
                gold[key + "_idx"] = idx
                gold[key + "_nseg"] = X.shape[-1]
                gold[key + "_X"] = X[..., idx]
        np.savez_compressed(os.path.join(GOLD, "spectra_%s.npz" % name), **gold)
        report.append(("spectra_" + name, "bit-exact vs reference"))

    protools_golden(report)
    hilbert_golden(report)
    analytic_golden(report)

    for name, status in report:
        print("%-18s %s" % (name, status))
    total = sum(os.path.getsize(os.path.join(GOLD, f)) for f in os.listdir(GOLD))
    print("golden bytes:", total)


if __name__ == "__main__":
    main()
