"""End-to-end throughput of BASELINE.json's configs 1-4 through the public
operator API (host chunks in, host results out), next to the CPU oracle on a
slice of the same workload.  The contract line of the round comes from bench.py
(config 5); this table shows every other config running at scale.

    python tools/configs_bench.py            # on a B200 box
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from openseize_b200 import producer  # noqa: E402
from openseize_b200.filtering.fir import Kaiser  # noqa: E402
from openseize_b200.filtering.iir import Butter  # noqa: E402
from openseize_b200.resampling.resampling import downsample  # noqa: E402
from openseize_b200.spectra.estimators import psd, stft  # noqa: E402


def cyclic_source(rows, chunk, nchunks, seed):
    """A generator producer over a pool of two pinned chunks (the recording is
    never materialised: SURVEY.md 8d)."""
    rng = np.random.default_rng(seed)
    pool = []
    for _ in range(2):
        t = torch.empty((rows, chunk), dtype=torch.float64, pin_memory=True)
        a = t.numpy()
        for r0 in range(0, rows, 32):
            a[r0:r0 + 32] = rng.standard_normal((min(32, rows - r0), chunk))
        pool.append(a)

    def gen():
        for i in range(nchunks):
            yield pool[i % 2]

    return producer(gen, chunk, -1, shape=(rows, chunk * nchunks)), pool


def drain(pro):
    n = 0
    for arr in pro:
        n += arr.shape[-1]
    torch.cuda.synchronize()
    return n


def timed(make, reps=2):
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        make()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best


def main():
    chunk = 1_000_000
    rows_out = []

    def report(name, rows, n, secs, cpu_rate, note):
        rate = rows * n / secs
        rows_out.append((name, rows, n, secs, rate, cpu_rate, note))
        print("%-44s %3d x %-10d %7.3f s  %8.2f G ch-samples/s   CPU oracle (1 core) %6.1f M/s   %s"
              % (name, rows, n, secs, rate / 1e9, cpu_rate / 1e6, note), flush=True)

    def cpu(fn, rows, n):
        x = np.random.default_rng(1).standard_normal((rows, n))
        t0 = time.perf_counter()
        fn(x)
        return rows * n / (time.perf_counter() - t0)

    # C1: Kaiser FIR 113 taps, 4 ch x 18M, 5 kHz
    k1 = Kaiser(fpass=500, fstop=600, fs=5000)
    src, _ = cyclic_source(4, chunk, 18, 1)
    t = timed(lambda: drain(k1(src, chunk, axis=-1, mode="same")))
    report("C1 Kaiser FIR 113 taps 'same'", 4, 18 * chunk, t,
           cpu(lambda x: oracle.oaconvolve(x, k1.coeffs, chunk, -1, "same"), 4, 2_000_000),
           "full size")
    # C2: Butterworth band-pass sosfiltfilt, 64 ch x 18M
    b2 = Butter(fpass=[1, 100], fstop=[0.5, 200], fs=5000, gpass=1, gstop=40)
    src, _ = cyclic_source(64, chunk, 18, 2)
    t = timed(lambda: drain(b2(src, chunk, axis=-1, dephase=True)))
    report("C2 Butterworth 8-section sosfiltfilt", 64, 18 * chunk, t,
           cpu(lambda x: oracle.sosfiltfilt(x, b2.coeffs, chunk, -1), 4, 2_000_000), "full size")
    # C3: downsample 5000 -> 250 Hz, 64 ch x 108M (slice: 24M)
    src, _ = cyclic_source(64, chunk, 24, 3)
    t = timed(lambda: drain(downsample(src, 20, 5000, chunk, axis=-1)))
    report("C3 polyphase downsample M=20 (449 taps)", 64, 24 * chunk, t,
           cpu(lambda x: oracle.polyphase_resample(x, 1, 20, 5000, chunk, -1), 4, 3_000_000),
           "24M of 108M samples")
    # C4: Welch PSD and STFT, 256 ch x 30 kHz (slice: 6M of 108M samples)
    src, _ = cyclic_source(256, chunk, 6, 4)
    t = timed(lambda: psd(src, 30000, axis=-1, resolution=30000 / 4096))
    report("C4 Welch PSD nfft 4096", 256, 6 * chunk, t,
           cpu(lambda x: oracle.welch_psd(x, 30000, -1, 30000 / 4096), 8, 1_000_000),
           "6M of 108M samples")
    src, _ = cyclic_source(32, chunk, 6, 5)

    def run_stft():
        f, tt, X = stft(src, 30000, axis=-1, resolution=30000 / 4096, asarray=False)
        drain(X)

    t = timed(run_stft)
    report("C4 STFT nfft 4096 (complex128 out, 32 ch)", 32, 6 * chunk, t,
           cpu(lambda x: oracle.stft(x, 30000, -1, 30000 / 4096), 4, 1_000_000),
           "one GPU's 32-channel share, 6M samples")
    print()
    print("| config | rows x samples | s | G ch-samples/s (e2e, host in / host out) | CPU oracle, "
          "1 core, M ch-samples/s | note |")
    print("|---|---|---|---|---|---|")
    for name, rows, n, secs, rate, cpu_rate, note in rows_out:
        print("| %s | %d x %d | %.3f | %.2f | %.1f | %s |" % (name, rows, n, secs, rate / 1e9,
                                                            cpu_rate / 1e6, note))


if __name__ == "__main__":
    main()
