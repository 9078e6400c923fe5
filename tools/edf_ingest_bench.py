"""End-to-end ingest of an EDF recording: int16 records decoded on the GPU
(ReaderProducer over file_io.edf.Reader) against the same samples handed over
as float64 host chunks.  Usage on a GPU box: python tools/edf_ingest_bench.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openseize_b200 import producer  # noqa: E402
from openseize_b200.file_io import edf  # noqa: E402
from openseize_b200.spectra.estimators import psd  # noqa: E402


def main():
    fs, nch, n = 5000, 64, 4_000_000
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((nch, n)) * 50).astype(np.float64)
    d = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    path = edf.write_edf(os.path.join(d, "osz_bench.edf"), x, fs, record_samples=fs)
    reader = edf.Reader(path)
    y = reader.read(0)

    def run(source):
        t0 = time.perf_counter()
        cnt, f, p = psd(source, fs, resolution=fs / 4096)
        torch.cuda.synchronize()
        return time.perf_counter() - t0, p

    class HostDecode:
        """The reference's way: Reader.read decodes to float64 on the host."""
        def __init__(self, r):
            self.r, self.shape = r, r.shape
        def read(self, a, b):
            return self.r.read(a, b)
        def open(self):
            self.r.open()
        def close(self):
            self.r.close()

    for _ in range(2):
        t_raw, p_raw = run(producer(edf.Reader(path), 1_000_000, -1))
        t_f64, p_f64 = run(producer(y, 1_000_000, -1))
    t_host, p_host = run(producer(HostDecode(edf.Reader(path)), 1_000_000, -1))
    print("EDF float64 decoded on the host: %.3f s  %.2f G ch-samples/s (Reader.read per chunk)"
          % (t_host, nch * n / t_host / 1e9))
    assert np.max(np.abs(p_raw - p_f64)) <= 1e-13 * np.max(np.abs(p_f64))
    print("EDF int16 -> GPU decode -> psd : %.3f s  %.2f G ch-samples/s (file read + %d MB over PCIe)"
          % (t_raw, nch * n / t_raw / 1e9, nch * n * 2 >> 20))
    print("float64 host chunks -> psd     : %.3f s  %.2f G ch-samples/s (%d MB over PCIe)"
          % (t_f64, nch * n / t_f64 / 1e9, nch * n * 8 >> 20))
    os.remove(path)


if __name__ == "__main__":
    main()
