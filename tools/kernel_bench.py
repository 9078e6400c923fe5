"""Per-kernel device timings at the BASELINE config sizes (CUDA events, inputs
larger than L2).  Scratch tool for tuning; the contract numbers come from
bench.py.  Usage on the GPU box: python tools/kernel_bench.py [name ...]"""

import json
import os
import sys

import numpy as np
import scipy.signal as sps
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openseize_b200.core import device as dv  # noqa: E402

PEAK = 6534.1
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]


ONCE = False


def timeit(fn, reps=5, warm=2):
    if ONCE:
        reps, warm = 1, 0
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    return float(np.median(ts))


def report(name, secs, ch_samples, bytes_per):
    gbs = ch_samples * bytes_per / secs / 1e9
    print("%-34s %8.3f ms  %8.2f Gsamp/s  %8.1f GB/s alg  %5.1f%% of %.0f" %
          (name, secs * 1e3, ch_samples / secs / 1e9, gbs, 100 * gbs / PEAK, PEAK), flush=True)


def main(which):
    dv.require_cuda()
    g = torch.Generator(device="cuda").manual_seed(0)

    def rnd(rows, n):
        return torch.randn((rows, n), dtype=torch.float64, device="cuda", generator=g)

    rows, n = 256, 1_000_000
    from openseize_b200.filtering.fir import Kaiser

    if not which or "fir" in which:
        for fs, label in ((5000, "113"), (30000, "671")):
            taps = Kaiser(500, 600, fs).coeffs
            x = rnd(rows, n + len(taps) - 1)
            y = torch.empty((rows, n), dtype=torch.float64, device="cuda")
            for algo, an in ((1, "direct"), (2, "fft"), (3, "fft float32 compute")):
                if algo == 1 and len(taps) > 200:
                    continue
                plan = dv.FirPlan(taps, algo)
                t = timeit(lambda: plan.run(x, n, out=y))
                report("fir %s taps %s" % (label, an), t, rows * n, 16)
            del x, y
    if not which or "sos" in which:
        sos = sps.butter(8, [1, 100], btype="bandpass", fs=5000, output="sos")
        plan = dv.SosPlan(sos)
        for r in (64, 256):
            x = rnd(r, n)
            y = torch.empty_like(x)
            st = dv.zeros((r, 8, 2))
            report("sos 8 sections fwd rows=%d" % r, timeit(lambda: plan.run(x, st, out=y)), r * n, 16)
            report("sos 8 sections state-only rows=%d" % r,
                   timeit(lambda: plan.run(x, st, want_output=False)), r * n, 8)
        b, a = sps.iirnotch(60, 10, fs=30000)
        plan = dv.SosPlan(np.concatenate([b, a])[None])
        for r in (256, 128, 64, 32):
            x = rnd(r, n)
            y = torch.empty_like(x)
            st = dv.zeros((r, 1, 2))
            report("notch biquad fwd rows=%d" % r, timeit(lambda: plan.run(x, st, out=y)), r * n, 16)
            report("notch biquad bwd rows=%d" % r,
                   timeit(lambda: plan.run(x, st, reverse=True, out=y)), r * n, 16)
            del x, y
    if "tf" in which:
        # (b, a) of order 8 (Butterworth band-pass): companion-matrix scan against the
        # sequential kernel
        b, a = sps.butter(4, [0.05, 0.3], btype="bandpass")
        plan = dv.TfPlan(b, a)
        for r in (256, 32):
            x = rnd(r, n)
            y = torch.empty_like(x)
            st = dv.zeros((r, 8))
            for mode in ("split", "scan", "seq"):
                os.environ["OSZ_TF_KERNEL"] = mode
                report("(b,a) order 8 %-5s rows=%d" % (mode, r),
                       timeit(lambda: plan.run(x, st, out=y), reps=3, warm=1), r * n, 16)
        os.environ.pop("OSZ_TF_KERNEL", None)
    if "sostile" in which:
        # one section: tiled look-back scan (default) against one CTA per row / time splits
        b, a = sps.iirnotch(60, 10, fs=30000)
        plan = dv.SosPlan(np.concatenate([b, a])[None])
        for r in (256, 128, 64, 32, 16, 8):
            x = rnd(r, n)
            y = torch.empty_like(x)
            st = dv.zeros((r, 1, 2))
            for mode in ("1", "1n", "0"):
                os.environ["OSZ_SOS_TILE"] = mode[0]
                os.environ["OSZ_SOS_TILE_TMA"] = "0" if mode == "1n" else "1"
                tag = {"1": "tma ", "1n": "tile", "0": "row "}[mode]
                report("notch %s fwd rows=%d" % (tag, r), timeit(lambda: plan.run(x, st, out=y)), r * n, 16)
                report("notch %s bwd rows=%d" % (tag, r),
                       timeit(lambda: plan.run(x, st, reverse=True, out=y)), r * n, 16)
                xs = x[:, :66_000]
                report("notch %s state-only 66k rows=%d" % (tag, r),
                       timeit(lambda: plan.run(xs, st, want_output=False)), r * 66_000, 8)
            del x, y
        os.environ.pop("OSZ_SOS_TILE", None)
    if not which or "upfirdn" in which:
        for fs, M in ((5000, 20), (30000, 25)):
            import oracle

            h = oracle.resample_filter(1, M, fs)
            x = rnd(rows, n)
            nout = n // M - 64
            for kern in ("mma", "polyphase"):
                plan = dv.UpfirdnPlan(h, 1, M, kernel=kern)
                report("downsample M=%d taps=%d %s" % (M, len(h), plan.kernel),
                       timeit(lambda: plan.run(x, 0, 32, nout)), rows * n, 8 * (1 + 1 / M))
            p32 = dv.UpfirdnPlan(h, 1, M, "float32")
            report("downsample M=%d taps=%d float32 compute" % (M, len(h)),
                   timeit(lambda: p32.run(x, 0, 32, nout)), rows * n, 8 * (1 + 1 / M))
            del x
        # the FIR(671) * anti-alias(561) cascade of config 5 as ONE decimating filter
        h = np.convolve(oracle.resample_filter(1, 25, 30000), Kaiser(500, 600, 30000).coeffs)
        x = rnd(rows, n)
        nout = n // 25 - 128
        for kern in ("mma", "polyphase"):
            plan = dv.UpfirdnPlan(h, 1, 25, kernel=kern)
            report("fused FIR+downsample M=25 taps=%d %s" % (len(h), plan.kernel),
                   timeit(lambda: plan.run(x, 0, 64, nout)), rows * n, 8 * (1 + 1 / 25))
        for r2 in (32, 64, 128):
            x2 = rnd(r2, n)
            plan = dv.UpfirdnPlan(h, 1, 25)
            report("fused FIR+downsample rows=%d %s" % (r2, plan.kernel),
                   timeit(lambda: plan.run(x2, 0, 64, nout)), r2 * n, 8 * (1 + 1 / 25))
            del x2
        p32 = dv.UpfirdnPlan(h, 1, 25, "float32")
        report("fused FIR+downsample M=25 taps=%d float32 compute" % len(h),
               timeit(lambda: p32.run(x, 0, 64, nout)), rows * n, 8 * (1 + 1 / 25))
        del x
    if "f32io" in which:
        # float32 I/O mode: float samples in and out (8 / 4.2 / 4 bytes per sample)
        dv.set_io("float32")
        try:
            import oracle

            xf = torch.randn((rows, n + 700), dtype=torch.float32, device="cuda", generator=g)
            for fs, label in ((5000, "113"), (30000, "671")):
                taps = Kaiser(500, 600, fs).coeffs
                plan = dv.FirPlan.cached(taps)
                yf = torch.empty((rows, n), dtype=torch.float32, device="cuda")
                report("fir %s taps float32 I/O" % label,
                       timeit(lambda: plan.run(xf, n, out=yf)), rows * n, 8)
            for fs, M in ((5000, 20), (30000, 25)):
                h = oracle.resample_filter(1, M, fs)
                plan = dv.UpfirdnPlan.cached(h, 1, M)
                nout = n // M - 64
                report("downsample M=%d taps=%d float32 I/O" % (M, len(h)),
                       timeit(lambda: plan.run(xf, 0, 32, nout)), rows * n, 4 * (1 + 1 / M))
            w = sps.get_window("hann", 4096)
            plan = dv.SpecPlan.cached(4096, 2048, w, "constant", 1.0 / (30000 * np.sum(w ** 2)))
            nseg = plan.nseg_available(n)
            acc = dv.zeros((rows, 2049))
            report("welch nfft=4096 float32 I/O", timeit(lambda: plan.welch_accum(xf, nseg, acc)),
                   rows * nseg * plan.stride, 4)
            b, a = sps.iirnotch(60, 10, fs=30000)
            plan = dv.SosPlan(np.concatenate([b, a])[None])
            st = dv.zeros((rows, 1, 2))
            yf = torch.empty((rows, n), dtype=torch.float32, device="cuda")
            report("notch biquad fwd float32 I/O", timeit(lambda: plan.run(xf[:, :n], st, out=yf)),
                   rows * n, 8)
        finally:
            dv.set_io("float64")
    if not which or "sosdec" in which:
        # backward notch pass + FIR(671) * anti-alias(561) decimator M=25 as one kernel
        import oracle

        h = np.convolve(oracle.resample_filter(1, 25, 30000), Kaiser(500, 600, 30000).coeffs)
        ufd = dv.UpfirdnPlan(h, 1, 25)
        b, a = sps.iirnotch(60, 10, fs=30000)
        sos = dv.SosPlan(np.concatenate([b, a])[None])
        for r2 in (256, 128, 64, 32):
            x2 = rnd(r2, n)
            st = dv.zeros((r2, 1, 2))
            nspan = dv.sosdec_spans(sos, ufd, r2, n)
            out = dv.empty((r2, n // 25))
            report("bwd notch + FIR + decimate rows=%d spans=%d" % (r2, nspan),
                   timeit(lambda: dv.sosdec_exec(sos, ufd, x2, True, st, nspan, 0, out, 0)),
                   r2 * n, 8 * (1 + 1 / 25))
            y2 = torch.empty_like(x2)
            report("  (separately: bwd notch pass rows=%d)" % r2,
                   timeit(lambda: sos.run(x2, st, reverse=True, out=y2)), r2 * n, 16)
            del x2, y2
    if not which or "welch" in which:
        for nfft in (1024, 4096, 8192) + ((2400, 10000, 60000) if "generic" in which else ()):
            w = sps.get_window("hann", nfft)
            plan = dv.SpecPlan(nfft, nfft // 2, w, "constant", 1.0 / (30000 * np.sum(w ** 2)))
            if nfft > 8192:
                rows = 64
            x = rnd(rows, n)
            nseg = plan.nseg_available(n)
            acc = dv.zeros((rows, nfft // 2 + 1))
            report("welch nfft=%d" % nfft, timeit(lambda: plan.welch_accum(x, nseg, acc)),
                   rows * nseg * plan.stride, 8)
            if nfft <= 4096 and not nfft & (nfft - 1):
                p32 = dv.SpecPlan(nfft, nfft // 2, w, "constant",
                                  1.0 / (30000 * np.sum(w ** 2)), "float32")
                report("welch nfft=%d float32 compute" % nfft,
                       timeit(lambda: p32.welch_accum(x, nseg, acc)), rows * nseg * plan.stride, 8)
            if nfft == 4096:
                xs = rnd(32, n)
                ns = plan.nseg_available(n)
                report("stft nfft=4096 rows=32", timeit(lambda: plan.segments(xs, ns, True)),
                       32 * ns * plan.stride, 24)
                report("stft nfft=4096 rows=32 float32 compute",
                       timeit(lambda: p32.segments(xs, ns, True)), 32 * ns * plan.stride, 24)
            del x


def protools_bench():
    """Producer tools: mask compaction, moments, standardize."""
    g = torch.Generator(device="cuda").manual_seed(1)
    rows, n = 256, 1_000_000
    x = torch.randn((rows, n), dtype=torch.float64, device="cuda", generator=g)
    mask = np.repeat(np.random.default_rng(0).random(n // 500) < 0.6, 500)
    idx = np.flatnonzero(mask)
    report("take_cols (60 %% kept, runs of 500)", timeit(lambda: dv.take_cols(x, idx)),
           rows * idx.size, 16)
    mom = dv.RowMoments(rows)
    report("row_moments", timeit(lambda: mom.add(x)), rows * n, 8)
    mu, sd = dv.zeros((rows,)), dv.zeros((rows,)) + 1.0
    y = torch.empty_like(x)
    report("row_standardize", timeit(lambda: dv.row_standardize(x, mu, sd, out=y)), rows * n, 16)
    report("col_moments standardize", timeit(lambda: dv.col_moments(x, True, "standardize")),
           rows * n, 16)


if __name__ == "__main__":
    argv = sys.argv[1:]
    if "--once" in argv:
        ONCE = True
        argv.remove("--once")
    if argv == ["protools"]:
        dv.require_cuda()
        protools_bench()
    else:
        main(argv)
