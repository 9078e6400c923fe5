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
    if size > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
