"""Time-sharded operators on real GPUs: run under torchrun (one rank per GPU,
NCCL), compare every rank-gathered result with the CPU oracle on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
        --master-addr 127.0.0.1 --master-port 29511 tools/shard_check.py
"""
import os
import sys

import numpy as np
import scipy.signal as sps
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank = int(os.environ.get("RANK", 0))
    size = int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    if size > 1:
        dist.init_process_group("nccl")
    import oracle
    from openseize_b200 import sharding
    from openseize_b200.filtering.fir import Kaiser

    fs = 5000
    rng = np.random.default_rng(17)
    x = rng.standard_normal((4, 600_000)) + 0.25
    taps = Kaiser(500, 600, fs).coeffs
    sos = sps.butter(8, [1, 100], btype="bandpass", fs=fs, output="sos")
    ba = sps.iirnotch(60, 10, fs=fs)
    got = dict(
        fir=sharding.fir_time_sharded(x, taps, 100_000, mode="same"),
        rs=sharding.resample_time_sharded(x, 1, 20, fs, 100_000),
        ff=sharding.iir_time_sharded(x, sos, 100_000, dephase=True),
        fw=sharding.iir_time_sharded(x, sos, 100_000, dephase=False),
        nf=sharding.iir_time_sharded(x, ba, 100_000, dephase=True, fmt="ba"),
    )
    cnt, f, p = sharding.psd_time_sharded(x, fs, resolution=fs / 4096)
    if rank == 0:
        cat = lambda blocks: np.concatenate(blocks, -1)
        ref = dict(
            fir=cat(oracle.oaconvolve(x, taps, 100_000, -1, "same")),
            rs=cat(oracle.polyphase_resample(x, 1, 20, fs, 100_000, -1)),
            ff=cat(oracle.sosfiltfilt(x, sos, 100_000, -1)),
            fw=cat(oracle.sosfilt(x, sos, 100_000, -1)[0]),
            nf=cat(oracle.filtfilt(x, ba, 100_000, -1)),
        )
        ok = True
        for k in ref:
            assert got[k].shape == ref[k].shape, (k, got[k].shape, ref[k].shape)
            err = np.max(np.abs(got[k] - ref[k])) / np.max(np.abs(ref[k]))
            ok &= err < 1e-9
            print("time-sharded %-4s world=%d  max err / peak = %.2e" % (k, size, err))
        rc, rf, rp = oracle.welch_psd(x, fs, -1, fs / 4096)
        err = np.max(np.abs(p - rp)) / np.max(np.abs(rp))
        ok &= err < 1e-9 and cnt == rc
        print("time-sharded psd  world=%d  max err / peak = %.2e  segments %d/%d" % (size, err, cnt, rc))
        print("SHARD_CHECK", "OK" if ok else "FAILED")
    # ---- BASELINE config 1 (4 channels x 18 M samples, Kaiser 113 taps, then Welch nfft
    #      4096): few channels, so the TIME axis is sharded.  Wall time of the sharded
    #      operators per rank count (host chunks in, gathered result out) and, separately,
    #      the one collective on the path: the all-reduce of the Welch partial sums.
    import json
    import time

    n1 = 18_000_000
    x1 = np.random.default_rng(3).standard_normal((4, n1))

    def wall(fn, reps=2):
        best = 1e9
        for _ in range(reps):
            if size > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            if size > 1:
                dist.barrier()
            best = min(best, time.perf_counter() - t0)
        return best

    t_fir = wall(lambda: sharding.fir_time_sharded(x1, taps, 1_000_000, mode="same", gather=False))
    t_psd = wall(lambda: sharding.psd_time_sharded(x1, fs, resolution=fs / 4096))
    msg = torch.zeros(4 * 2049 + 1, dtype=torch.float64, device="cuda")
    t_ar = None
    if size > 1:
        for _ in range(3):
            dist.all_reduce(msg)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            dist.all_reduce(msg)
        b.record()
        torch.cuda.synchronize()
        t_ar = a.elapsed_time(b) / 20
    # the same FIR shard from a READER source: every rank reads only its span + halo
    class _Reader:
        def __init__(self, arr):
            self.arr, self.read_samples = arr, 0
            self.shape = arr.shape

        def read(self, start, stop):
            self.read_samples += stop - start
            return self.arr[:, start:stop]

        def open(self):
            pass

        def close(self):
            pass

    from openseize_b200 import producer

    rd = _Reader(x1[:, :6_000_000])
    (o0, o1), loc = sharding.fir_time_sharded(producer(rd, 1_000_000, -1), taps, 1_000_000,
                                              mode="same", gather=False)
    refl = cat_ref = None
    if rank == 0:
        refl = np.concatenate(oracle.oaconvolve(x1[:, :6_000_000], taps, 1_000_000, -1, "same"), -1)
    full = sharding.gather_time(loc, -1)
    if rank == 0:
        err = np.max(np.abs(full - refl)) / np.max(np.abs(refl))
        print("time-sharded fir from a reader source world=%d  max err / peak = %.2e, rank 0 read "
              "%d of %d samples" % (size, err, rd.read_samples, 6_000_000))
        print("SHARD_BENCH", json.dumps({
            "world": size, "config": "C1: 4 ch x 18e6 float64 host array, Kaiser 113 taps 'same', "
                                     "psd nfft 4096",
            "fir_time_sharded_s": t_fir, "fir_channel_samples_per_s": 4 * n1 / t_fir,
            "psd_time_sharded_s": t_psd, "psd_channel_samples_per_s": 4 * n1 / t_psd,
            "welch_allreduce_ms": t_ar, "allreduce_bytes": int(msg.numel() * 8)}))
    if size > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
