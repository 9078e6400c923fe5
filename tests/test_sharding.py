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
        if rank == 0:
            out.put((cnt, f, p, c2, np.concatenate(gathered, 0)))
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
    cnt, f, p, c2, p2 = out.get(timeout=240)
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
