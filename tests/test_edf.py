"""EDF ingest: the reader mirrors the reference's
file_io/edf.py Reader; a ReaderProducer over it feeds the GPU path int16 records
that are calibrated on the device, bit-identically to Reader.read."""

import os
import pickle
import sys

import numpy as np
import pytest

from openseize_b200 import producer
from openseize_b200.file_io import edf
from tests.conftest import has_cuda

REF = "/root/reference/src"
GOLD = os.path.join(os.path.dirname(__file__), "golden", "edf_small.npz")


def _recording(rng, nch=5, n=6000):
    t = np.arange(n) / 500.0
    x = 40 * np.sin(2 * np.pi * 8 * t)[None] * rng.uniform(0.5, 2, (nch, 1))
    return x + 15 * rng.standard_normal((nch, n)) + rng.uniform(-30, 30, (nch, 1))


def test_reader_roundtrip(tmp_path):
    rng = np.random.default_rng(7)
    x = _recording(rng)
    path = edf.write_edf(tmp_path / "a.edf", x, fs=500, record_samples=250)
    with edf.Reader(path) as r:
        assert r.shape == x.shape and r.header.num_records == 24
        y = r.read(0)
        # quantisation to int16 over each channel's physical range
        step = r.header.slopes
        assert np.all(np.abs(y - x) <= 0.5 * step[:, None] + 1e-9)
        assert np.array_equal(r.read(123, 4321), y[:, 123:4321])
        raw = r.read_raw(123, 4321)
        assert raw.records.dtype == np.int16 and np.array_equal(raw.decode(), y[:, 123:4321])
        r.channels = [3, 1]
        assert np.array_equal(r.read(10, 700), y[[3, 1], 10:700])
        assert np.array_equal(r.read_raw(10, 700).decode(), y[[3, 1], 10:700])
        assert r.read(10 ** 7).shape == (2, 0)
    r2 = pickle.loads(pickle.dumps(edf.Reader(path)))
    assert np.array_equal(r2.read(0, 50), y[:, :50])
    r2.close()


def test_golden_file(tmp_path):
    """A small EDF written here must decode to the values the REFERENCE reader
    returned for the same bytes (tests/golden/edf_small.npz, made by
    test_against_reference below when /root/reference is present)."""
    g = np.load(GOLD)
    path = tmp_path / "g.edf"
    path.write_bytes(g["file_bytes"].tobytes())
    with edf.Reader(path) as r:
        assert r.shape == tuple(g["shape"])
        assert np.array_equal(r.read(0), g["all"])
        assert np.array_equal(r.read(77, 1901), g["span"])
        assert np.array_equal(r.header.slopes, g["slopes"])
        assert np.array_equal(r.header.offsets, g["offsets"])


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference not present on this box")
def test_against_reference(tmp_path):
    sys.path.insert(0, REF)
    try:
        from openseize.file_io import edf as ref_edf
    finally:
        sys.path.remove(REF)
    rng = np.random.default_rng(3)
    x = _recording(rng, nch=4, n=3000)
    path = edf.write_edf(tmp_path / "r.edf", x, fs=500, record_samples=300)
    with ref_edf.Reader(path) as rr, edf.Reader(path) as mine:
        assert rr.shape == mine.shape
        for key in ("num_records", "samples_per_record", "names", "physical_min", "digital_max"):
            assert rr.header[key] == mine.header[key], key
        assert np.array_equal(rr.header.slopes, mine.header.slopes)
        for a, b in ((0, None), (5, 299), (300, 301), (299, 2999), (2500, 5000)):
            assert np.array_equal(rr.read(a, b), mine.read(a, b)), (a, b)
        full, span = rr.read(0), rr.read(77, 1901)
        slopes, offsets = np.array(rr.header.slopes), np.array(rr.header.offsets)
    if not os.path.exists(GOLD):
        np.savez_compressed(GOLD, file_bytes=np.frombuffer(open(path, "rb").read(), dtype=np.uint8),
                            shape=np.array(full.shape), all=full, span=span, slopes=slopes,
                            offsets=offsets)


def test_reader_producer_host_and_raw(tmp_path, fake_gpu):
    """Host iteration yields the reference's float64 chunks; the device path
    takes RawChunks (decoded by the stand-in here, by the kernel on a GPU)."""
    from openseize_b200.filtering.fir import Kaiser

    rng = np.random.default_rng(11)
    x = _recording(rng, nch=3, n=8000)
    path = edf.write_edf(tmp_path / "p.edf", x, fs=500, record_samples=500)
    reader = edf.Reader(path)
    y = reader.read(0)
    pro = producer(reader, 1700, -1)
    assert np.array_equal(np.concatenate(list(pro), -1), y)
    chunks = list(pro.iter_raw())
    assert [c.shape[1] for c in chunks] == [1700, 1700, 1700, 1700, 1200]
    filt = Kaiser(50, 80, 500)
    got = filt(pro, 1700, axis=-1).to_array()
    ref = filt(producer(y, 1700, -1), 1700, axis=-1).to_array()
    assert np.array_equal(got, ref)
    assert pickle.loads(pickle.dumps(pro)).shape == pro.shape


@pytest.mark.gpu
def test_device_decode_is_bit_identical(tmp_path):
    if not has_cuda():
        pytest.skip("needs a GPU")
    from openseize_b200.core import device as dv
    from openseize_b200.core import numerical as nm
    from openseize_b200.spectra.estimators import psd

    rng = np.random.default_rng(5)
    x = _recording(rng, nch=6, n=40000)
    path = edf.write_edf(tmp_path / "d.edf", x, fs=500, record_samples=500)
    reader = edf.Reader(path)
    y = reader.read(0)
    pro = producer(reader, 7000, -1)
    blocks = [b.cpu().numpy() for b in nm.device_chunks(pro, -1, regrid=False)]
    assert np.array_equal(np.concatenate(blocks, -1), y)          # bit-identical calibration
    cnt, f, p = psd(producer(edf.Reader(path), 7000, -1), 500, resolution=500 / 1024)
    rc, rf, rp = psd(y, 500, resolution=500 / 1024)
    # (the Welch sum uses atomics: equal to rounding, not bitwise)
    assert cnt == rc and np.max(np.abs(p - rp)) <= 1e-13 * np.max(np.abs(rp))
