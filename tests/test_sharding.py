"""Multi-GPU host logic on CPU: world_size-2 gloo processes, kernels replaced
by the numpy stand-ins.  Covers the shard planning, that channel shards
reproduce the single-process result, and the one collective on the path (the
all-reduce of the time-sharded Welch sum)."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from openseize_b200 import sharding


def test_split_range_and_spans():
    assert sharding.split_range(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert sharding.split_range(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    n, nfft, ov = 100_000, 1024, 0.5
    nseg, stride = sharding.welch_segments(n, nfft, ov)
    assert (nseg, stride) == ((n - nfft) // 512 + 1, 512)
    covered = 0
    for size in (1, 2, 3, 8):
        spans = [sharding.time_span(n, nfft, ov, r, size) for r in range(size)]
        counts = [sharding.welch_segments(b - a, nfft, ov)[0] if b > a else 0 for a, b in spans]
        assert sum(counts) == nseg                      # every segment exactly once
        for (a, b), (c, d) in zip(spans, spans[1:]):
            assert c == b - nfft + stride               # next span starts one stride on
        covered += 1
    assert covered == 4
    assert sharding.time_span(500, 1024, 0.5, 0, 2) == (0, 0)
    index, split = sharding.channel_block((256, 1000), -1, 3, 8)
    assert split == 0 and index[0] == slice(96, 128)
    index, split = sharding.channel_block((1000, 6, 2), 0, 1, 2)
    assert split == 1 and index[1] == slice(3, 6)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, size, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        from tests import fake_backend

        fake_backend.install_raw()
        from openseize_b200 import producer
        from openseize_b200.filtering.fir import Kaiser
        from openseize_b200.spectra.estimators import psd

        rng = np.random.default_rng(5)
        x = rng.standard_normal((6, 60000)) + 1.0
        fs = 1024
        # (1) time-sharded Welch: one all-reduce
        cnt, f, p = sharding.psd_time_sharded(x, fs, resolution=1.0)
        # (2) channel-sharded FIR -> psd: no collective on the data path
        mine = sharding.shard_channels(x, 9000)
        filt = Kaiser(100, 150, fs)
        c2, f2, p2 = psd(filt(mine, 9000), fs, resolution=1.0)
        gathered = [None] * size
        dist.all_gather_object(gathered, p2)
        # (3) time-sharded filters: halos from the host array; the IIR state crosses
        #     ranks as an all-gather of (rows, nsec, 2) summaries
        import scipy.signal as sps

        xt = rng.standard_normal((3, 50000)) + 0.5
        fir_same = sharding.fir_time_sharded(xt, filt.coeffs, 7000, mode="same")
        fir_valid = sharding.fir_time_sharded(xt, filt.coeffs, 7000, mode="valid")
        rs = sharding.resample_time_sharded(xt, 3, 7, fs, 6000)
        sos = sps.butter(4, [5, 60], btype="bandpass", fs=fs, output="sos")
        ff = sharding.iir_time_sharded(xt, sos, 8000, dephase=True)
        fw = sharding.iir_time_sharded(xt, sos, 8000, dephase=False)
        ba = sps.iirnotch(60, 10, fs=fs)
        nf = sharding.iir_time_sharded(xt, ba, 8000, dephase=True, fmt="ba")
        if rank == 0:
            out.put((cnt, f, p, c2, np.concatenate(gathered, 0),
                     dict(fir_same=fir_same, fir_valid=fir_valid, rs=rs, ff=ff, fw=fw, nf=nf)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_gloo():
    import oracle

    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    cnt, f, p, c2, p2, sharded = out.get(timeout=240)
    for pr in procs:
        pr.join(60)
        assert pr.exitcode == 0
    rng = np.random.default_rng(5)
    x = rng.standard_normal((6, 60000)) + 1.0
    rc, rf, rp = oracle.welch_psd(x, 1024, -1, 1.0)
    assert cnt == rc and np.array_equal(f, rf)
    assert np.max(np.abs(p - rp)) / np.max(np.abs(rp)) < 1e-12
    from oracle.chunked import _kaiser_lowpass

    taps = _kaiser_lowpass(100, 150, 1024, 1.0, 40.0)
    y = np.concatenate(oracle.oaconvolve(x, taps, 9000, -1, "same"), -1)
    rc2, _, rp2 = oracle.welch_psd(y, 1024, -1, 1.0)
    assert c2 == rc2
    assert np.max(np.abs(p2 - rp2)) / np.max(np.abs(rp2)) < 1e-12

    # time-sharded filters against the single-process oracle
    import scipy.signal as sps

    xt = rng.standard_normal((3, 50000)) + 0.5
    for mode in ("same", "valid"):
        ref = np.concatenate(oracle.oaconvolve(xt, taps, 7000, -1, mode), -1)
        got = sharded["fir_" + mode]
        assert got.shape == ref.shape
        assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < 1e-12
    ref = np.concatenate(oracle.polyphase_resample(xt, 3, 7, 1024, 6000, -1), -1)
    assert sharded["rs"].shape == ref.shape
    assert np.max(np.abs(sharded["rs"] - ref)) / np.max(np.abs(ref)) < 1e-12
    sos = sps.butter(4, [5, 60], btype="bandpass", fs=1024, output="sos")
    ref = np.concatenate(oracle.sosfiltfilt(xt, sos, 8000, -1), -1)
    assert np.max(np.abs(sharded["ff"] - ref)) / np.max(np.abs(ref)) < 1e-9
    ref = np.concatenate(oracle.sosfilt(xt, sos, 8000, -1)[0], -1)
    assert np.max(np.abs(sharded["fw"] - ref)) / np.max(np.abs(ref)) < 1e-9
    ref = np.concatenate(oracle.filtfilt(xt, sps.iirnotch(60, 10, fs=1024), 8000, -1), -1)
    assert np.max(np.abs(sharded["nf"] - ref)) / np.max(np.abs(ref)) < 1e-9


def test_time_shards_simulated(fake_gpu, monkeypatch):
    """FIR and resampling time shards need no collective: emulate 1, 3 and 5
    ranks in one process and stitch the spans."""
    import oracle
    from oracle.chunked import _kaiser_lowpass

    rng = np.random.default_rng(11)
    x = rng.standard_normal((2, 3, 41000))
    taps = _kaiser_lowpass(100, 150, 1024, 1.0, 40.0)
    for size in (1, 3, 5):
        for mode in ("full", "same", "valid"):
            parts = []
            for r in range(size):
                monkeypatch.setattr(sharding, "world", lambda group=None, r=r: (r, size))
                (o0, o1), loc = sharding.fir_time_sharded(x, taps, 5000, mode=mode, gather=False)
                assert loc.shape[-1] == o1 - o0
                parts.append(loc)
            got = np.concatenate(parts, -1)
            ref = np.concatenate(oracle.oaconvolve(x, taps, 5000, -1, mode), -1)
            assert got.shape == ref.shape
            assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < 1e-12
        for (L, M) in ((1, 4), (3, 2), (5, 3)):
            parts = []
            for r in range(size):
                monkeypatch.setattr(sharding, "world", lambda group=None, r=r: (r, size))
                (o0, o1), loc = sharding.resample_time_sharded(x, L, M, 1024, 4000, gather=False)
                parts.append(loc)
            got = np.concatenate(parts, -1)
            ref = np.concatenate(oracle.polyphase_resample(x, L, M, 1024, 4000, -1), -1)
            assert got.shape == ref.shape, (got.shape, ref.shape, L, M, size)
            assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < 1e-12


class _ArrayReader:
    """Minimal reader (shape, read(start, stop), open, close): what ReaderProducer needs;
    `channels` restricts the rows like the EDF readers' attribute of that name."""

    def __init__(self, x):
        self.x, self.channels, self.reads = x, list(range(x.shape[0])), []

    @property
    def shape(self):
        return (len(self.channels), self.x.shape[1])

    def read(self, start, stop):
        self.reads.append((start, stop))
        return self.x[self.channels, start:stop]

    def open(self):
        pass

    def close(self):
        pass


def _gen_source(x, step):
    for i in range(0, x.shape[-1], step):
        yield x[..., i:i + step]


def test_shards_of_out_of_core_sources(fake_gpu, monkeypatch):
    """Channel and time shards of reader and generating-function producers (the
    out-of-core recording of BASELINE config 5): a rank reads only its channel block /
    its time span plus the filter halo, and the stitched result equals the
    single-process one."""
    import functools

    import oracle
    from openseize_b200 import producer
    from openseize_b200.filtering.fir import Kaiser
    from openseize_b200.spectra.estimators import psd
    from oracle.chunked import _kaiser_lowpass

    rng = np.random.default_rng(12)
    x = rng.standard_normal((6, 45000)) + 0.3
    fs = 1024
    filt = Kaiser(100, 150, fs)
    taps = _kaiser_lowpass(100, 150, fs, 1.0, 40.0)
    ref_fir = np.concatenate(oracle.oaconvolve(x, taps, 6000, -1, "same"), -1)
    rc, rf, ref_psd = oracle.welch_psd(ref_fir, fs, -1, 1.0)
    size = 3
    for kind in ("reader", "generator"):
        parts, readers = [], []
        for r in range(size):
            monkeypatch.setattr(sharding, "world", lambda group=None, r=r: (r, size))
            if kind == "reader":
                rd = _ArrayReader(x)
                src = producer(rd, 6000, -1)
                readers.append(rd)
            else:
                src = producer(functools.partial(_gen_source, x, 5000), 6000, -1, shape=x.shape)
            mine = sharding.shard_channels(src, 6000)
            assert mine.shape == (2, 45000)
            cnt, f, p = psd(filt(mine, 6000), fs, resolution=1.0)
            assert cnt == rc
            parts.append(p)
        got = np.concatenate(parts, 0)
        assert np.max(np.abs(got - ref_psd)) / np.max(np.abs(ref_psd)) < 1e-12
    # time shards of a reader: every rank reads its own span (+ halo) only
    parts, spans = [], []
    for r in range(size):
        monkeypatch.setattr(sharding, "world", lambda group=None, r=r: (r, size))
        rd = _ArrayReader(x)
        (o0, o1), loc = sharding.fir_time_sharded(producer(rd, 6000, -1), taps, 6000, mode="same",
                                                  gather=False)
        parts.append(loc)
        spans.append((min(a for a, _ in rd.reads), max(b for _, b in rd.reads)))
        assert spans[-1][1] - spans[-1][0] < x.shape[1] // size + 2 * len(taps)
    got = np.concatenate(parts, -1)
    assert np.max(np.abs(got - ref_fir)) / np.max(np.abs(ref_fir)) < 1e-12
    # ... and of a generating function (chunks outside the span are skipped)
    parts = []
    for r in range(size):
        monkeypatch.setattr(sharding, "world", lambda group=None, r=r: (r, size))
        src = producer(functools.partial(_gen_source, x, 5000), 6000, -1, shape=x.shape)
        (o0, o1), loc = sharding.resample_time_sharded(src, 1, 4, fs, 6000, gather=False)
        parts.append(loc)
    ref = np.concatenate(oracle.polyphase_resample(x, 1, 4, fs, 6000, -1), -1)
    got = np.concatenate(parts, -1)
    assert got.shape == ref.shape and np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < 1e-12


def test_cascade_transition_matches_recurrence():
    import scipy.signal as sps

    sos = sps.butter(3, [5, 60], btype="bandpass", fs=1024, output="sos")
    T = sharding.cascade_transition(sos).astype(np.float64)
    rng = np.random.default_rng(3)
    z = rng.standard_normal((sos.shape[0], 2))
    n = 37
    _, zf = sps.sosfilt(sos, np.zeros(n), zi=z)
    got = (np.linalg.matrix_power(T, n) @ z.reshape(-1)).reshape(-1, 2)
    assert np.allclose(got, zf, rtol=1e-10, atol=1e-14)
