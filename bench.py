"""Headline benchmark: channel-samples/sec through the filtering + PSD hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[4], the configuration `metric` is quoted on):
the full pipeline  Notch(60 Hz) forward-backward IIR -> Kaiser(500/600 Hz) FIR
(671 taps, mode 'same') -> downsample by 25 -> Welch PSD (nfft 4096, Hann, 50 %
overlap)  on a synthetic 256-channel x 30 kHz float64 recording streamed in
chunks of 1e6 samples.  The 24 h recording (2592 chunks, 5.3 TB) is an
out-of-core stream; a "step" is 4 chunks (4 x 256 x 1e6 channel-samples) of the
WHOLE recording passing through all four stages in steady state, and the
reported rate is what a full 24 h pass sustains.  One process per GPU; at N > 1
the recording's 256 channels are split into blocks of 256 / N per GPU (channel
sharding, no data-path collective, STRONG scaling: the job is the same at every
N); `weak_scaling` additionally reports N independent 256-channel recordings.

  value    : inputs already resident in HBM (a cyclic pool of device chunks),
             CUDA-event timed, max over ranks.
  e2e      : the same pipeline through the public producer/operator API with
             HOST (pinned) chunks; H2D of every chunk and D2H of the PSD inside
             the timed region; `h2d_roof_GBps` = what plain pinned copies reach
             on all ranks at once.
  parity   : the measured pipeline against the CPU oracle on channels of the
             same pool (outside the timed region).
  roofline : the kernel with the largest share of the step -- algorithmic bytes
             per launch / its CUDA-event time; `stages` counts a
             forward-backward filter as ONE read + ONE write per sample;
             `roofline_named` are the two kernels north_star names, timed alone.
  cpu_baseline / --impl reference : the UNMODIFIED reference (baseline/_ref)
             through its own operator API on all host cores (the oracle port
             when the reference is not installed).
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FS = 30000
ROWS = 256
CHUNK = 1_000_000
M_DEC = 25
NFFT = 4096
CPS_DEFAULT = 4          # chunks per step: a step is 4 x (256 x 1e6) channel-samples
METRIC = "channel-samples/sec filtered+PSD"
WORKLOAD = ("C5 pipeline: Notch(60,w6) filtfilt -> Kaiser(500,600) 671-tap FIR 'same' -> "
            "downsample M=25 -> Welch PSD nfft=4096 hann 50%; 256 ch x 30 kHz float64, "
            "chunksize 1e6, step = 4 chunks of the 24 h stream (2592 chunks)")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=ROWS)
    ap.add_argument("--chunk", type=int, default=CHUNK)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-narrow", action="store_true", help="skip the float32 / int16 e2e extras")
    ap.add_argument("--no-f32", action="store_true", help="skip the float32-compute extra pass")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the pool")
    ap.add_argument("--no-weak", action="store_true",
                    help="N > 1: skip the extra weak-scaling pass (256 channels per GPU)")
    ap.add_argument("--chunks-per-step", type=int, default=CPS_DEFAULT,
                    help="chunks of the recording per step")
    ap.add_argument("--no-named", action="store_true",
                    help="skip the stand-alone FIR / Welch kernel timings")
    ap.add_argument("--shard", default="channel", choices=["channel", "time"],
                    help="channel: the contract's run (config 5, channel blocks per GPU); "
                         "time: BASELINE config 1 (4 ch x 18e6) with the TIME axis sharded")
    return ap.parse_args()


def run_time_shard(args):
    """`--shard time`: BASELINE config 1 -- a few-channel recording (4 ch x 18e6 float64 host
    array; Kaiser 500/600 Hz at fs 5000, 113 taps, 'same'; Welch nfft 4096) whose TIME axis is
    split across the ranks: `fir_time_sharded` (every rank filters its span plus the filter
    halo) and `psd_time_sharded` (Welch segments split; ONE all-reduce of the partial sums),
    plus that all-reduce timed on its own.  Wall time per operator, max over ranks (they end
    on a barrier), host array in and results out: one JSON line from rank 0."""
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl")
    from openseize_b200 import sharding
    from openseize_b200.filtering.fir import Kaiser

    fs, n = 5000, 18_000_000
    x = np.random.default_rng(3).standard_normal((4, n))
    taps = Kaiser(500, 600, fs).coeffs

    def wall(fn, reps):
        best = float("inf")
        for _ in range(reps):
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            best = min(best, time.perf_counter() - t0)
        return best

    reps = max(2, min(args.steps, 5))
    for _ in range(max(1, min(args.warmup, 2))):
        sharding.fir_time_sharded(x, taps, 1_000_000, mode="same", gather=False)
        sharding.psd_time_sharded(x, fs, resolution=fs / 4096)
    t_fir = wall(lambda: sharding.fir_time_sharded(x, taps, 1_000_000, mode="same", gather=False), reps)
    t_psd = wall(lambda: sharding.psd_time_sharded(x, fs, resolution=fs / 4096), reps)
    t_ar = None
    if dist is not None:
        msg = torch.zeros(4 * 2049 + 1, dtype=torch.float64, device="cuda")
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
    if rank == 0:
        print(json.dumps({
            "metric": "channel-samples/sec, time-sharded FIR (config 1)", "value": 4 * n / t_fir,
            "unit": "channel-samples/s", "n_gpus": world, "higher_is_better": True,
            "scaling": "strong", "dtype": "f64", "data": "synthetic",
            "config": {"workload": "BASELINE config 1: 4 ch x 18e6 float64 host array, Kaiser "
                                   "500/600 Hz @ 5 kHz (113 taps) oaconvolve 'same', chunksize 1e6; "
                                   "Welch PSD nfft 4096; time axis split across the ranks",
                       "sharding": "time spans with filter-length halos; one all-reduce of the "
                                   "Welch partial sums"},
            "fir_time_sharded_s": t_fir, "psd_time_sharded_s": t_psd,
            "psd_channel_samples_per_s": 4 * n / t_psd,
            "welch_allreduce_ms": t_ar, "allreduce_bytes": 8 * (4 * 2049 + 1),
            "note": "wall time, pageable host array in and results out: bound by the box's "
                    "host-to-device bandwidth (the kernels take 0.4 ms / 0.35 ms)"}))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# ---------------------------------------------------------------------------
# the pipeline, through the public operator API
# ---------------------------------------------------------------------------
def build_pipeline(source, chunk):
    from openseize_b200.filtering.fir import Kaiser
    from openseize_b200.filtering.iir import Notch
    from openseize_b200.resampling.resampling import downsample

    notch = Notch(fstop=60, width=6, fs=FS)
    kais = Kaiser(fpass=500, fstop=600, fs=FS)
    assert len(kais.coeffs) == 671
    p1 = notch(source, chunk, axis=-1, dephase=True)
    p2 = kais(p1, chunk, axis=-1, mode="same")
    return downsample(p2, M_DEC, FS, chunk, axis=-1)


def run_psd(pro):
    from openseize_b200.spectra.estimators import psd

    fs2 = FS // M_DEC
    return psd(pro, fs2, axis=-1, resolution=fs2 / NFFT)


class Marks:
    """The source calls `at(i)` before handing out chunk i: at chunk `warmup`
    and `warmup + steps` the device is drained, ranks meet at a barrier and a
    CUDA event is recorded -- the timed region holds exactly `steps` chunks of
    steady-state work."""

    def __init__(self, warmup, steps, barrier):
        import torch

        self.torch, self.w, self.k, self.barrier = torch, warmup, steps, barrier
        self.ev = {}
        self.wall = {}
        self.host = []           # host clock at every chunk hand-out (diagnostic)
        self.on_start, self.on_stop = None, None

    def host_pace(self):
        """How fast the host enqueues: ms between consecutive chunk hand-outs inside
        the timed region.  A median near the device's ms per step means the host
        cannot run ahead of the GPU and every host hiccup lands in the step time."""
        t = [b - a for (i, a), (j, b) in zip(self.host, self.host[1:])
             if self.w < i and j < self.w + self.k]
        if not t:
            return None
        return {"median_ms": round(1e3 * float(np.median(t)), 3),
                "max_ms": round(1e3 * float(np.max(t)), 3)}

    def at(self, i):
        self.host.append((i, time.perf_counter()))
        if i not in (self.w, self.w + self.k):
            return
        t = self.torch
        if i == self.w + self.k:
            # this rank's own finish, before it waits for the others at the barrier
            self.own_stop = t.cuda.Event(enable_timing=True)
            self.own_stop.record()
        t.cuda.synchronize()
        self.barrier()
        if i == self.w and self.on_start:
            self.on_start()
        ev = t.cuda.Event(enable_timing=True)
        ev.record()
        t.cuda.synchronize()
        self.ev[i], self.wall[i] = ev, time.perf_counter()
        if i == self.w + self.k and self.on_stop:
            self.on_stop()

    def seconds(self):
        return self.ev[self.w].elapsed_time(self.ev[self.w + self.k]) * 1e-3

    def own_seconds(self):
        return self.ev[self.w].elapsed_time(self.own_stop) * 1e-3


TAIL = 3   # untimed cool-down chunks so the pipeline's look-ahead never drains inside the timing


def device_source(pool, rows, chunk, nchunks, marks):
    from openseize_b200.core.producer import DeviceProducer

    def gen():
        for i in range(nchunks):
            marks.at(i)
            yield pool[i % len(pool)]

    return DeviceProducer(gen, chunk, (rows, chunk * nchunks))


class DeviceSourceOnce:
    """A short device-resident recording without marks (pre-warm)."""

    def __init__(self, pool, rows, chunk, nchunks):
        self.pool, self.rows, self.chunk, self.nchunks = pool, rows, chunk, nchunks

    def producer(self):
        from openseize_b200.core.producer import DeviceProducer

        def gen():
            for i in range(self.nchunks):
                yield self.pool[i % len(self.pool)]

        return DeviceProducer(gen, self.chunk, (self.rows, self.chunk * self.nchunks))


def host_source(pool, rows, chunk, nchunks, marks):
    from openseize_b200 import producer

    def gen():
        for i in range(nchunks):
            marks.at(i)
            yield pool[i % len(pool)]

    return producer(gen, chunk, -1, shape=(rows, chunk * nchunks))


class ClockSampler:
    """SM clock and throttle reasons while the timed region runs, read through NVML
    in a background thread (the same counters `nvidia-smi --query-gpu=clocks.sm,
    clocks_event_reasons.*` prints).  An `nvidia-smi -lms 20` subprocess was used
    first: every one of its samples stalled the GPU work for ~3 ms (64.5 instead of
    74.5 G channel-samples/s over a 140 ms timed region), eight of them at N = 8
    stretched the step from 3.5 to 5.6 ms, and its ~0.5 s start-up stalls launches
    too.  In-process NVML queries cost microseconds."""

    REASONS = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown", 0x8),
               ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
               ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
               ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap", 0x4))

    def __init__(self, index, enabled=True, period_ms=20):
        self.index, self.enabled, self.period = index, enabled, period_ms / 1e3
        self.samples, self.thread, self._stop = [], None, threading.Event()
        self.t0 = self.t1 = None      # wall-clock bounds of the timed region
        self.ready = False

    def mark_start(self):
        self.t0 = time.time()

    def mark_stop(self):
        self.t1 = time.time()

    def start(self):
        if not self.enabled:
            return
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            # LOCAL_RANK indexes CUDA devices; honour CUDA_VISIBLE_DEVICES
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = self.index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if self.index < len(ids) and ids[self.index].isdigit():
                    phys = int(ids[self.index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle,
                                                                 pynvml.NVML_CLOCK_SM))
        except Exception:
            self.enabled = False
            return
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def _poll(self):
        n = self.nvml
        while not self._stop.is_set():
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                try:
                    mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.samples.append((time.time(), sm, mask))
                self.ready = True
            except Exception:
                pass
            self._stop.wait(self.period)

    def stop(self):
        self._stop.set()
        if self.thread:
            self.thread.join(timeout=2)

    def summary(self):
        sm, reasons = [], set()
        for stamp, clock, mask in self.samples:
            if self.t0 is not None and self.t1 is not None and not (
                    self.t0 - 0.02 <= stamp <= self.t1 + 0.02):
                continue
            sm.append(clock)
            for name, _, bit in self.REASONS:
                if mask & bit:
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": self.max_sm,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvml"}


# ---------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path, all host cores
# ---------------------------------------------------------------------------
def _cpu_worker(args):
    seed, rows, n, chunk = args
    import scipy.signal as sps

    import oracle

    rng = np.random.default_rng([0, seed])
    x = rng.standard_normal((rows, n))
    b, a = sps.iirnotch(60, 60 / 6, fs=FS)
    t0 = time.perf_counter()
    r1 = np.concatenate(oracle.filtfilt(x, (b, a), chunk, -1), -1)
    from oracle.chunked import _kaiser_lowpass

    taps = _kaiser_lowpass(500, 600, FS, 1.0, 40.0)
    r2 = np.concatenate(oracle.oaconvolve(r1, taps, chunk, -1, "same"), -1)
    r3 = np.concatenate(oracle.polyphase_resample(r2, 1, M_DEC, FS, chunk, -1), -1)
    fs2 = FS // M_DEC
    cnt, _, p = oracle.welch_psd(r3, fs2, -1, fs2 / NFFT)
    return time.perf_counter() - t0, float(p.sum()), cnt


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------
# CPU arm, preferred: the UNMODIFIED reference (baseline/_ref, pure Python over
# numpy/scipy) driven through its own public operator API at chunksize 1e6
# ---------------------------------------------------------------------------
REF_CHUNKS = 4          # chunks of 1e6 samples per worker and step (the resampler needs >= 3)


def _ref_worker(args):
    """One worker = one channel block of the recording through the reference's
    Notch -> Kaiser -> downsample -> psd, exactly as a user of openseize writes it
    (producer of a generating function, chunksize 1e6, lazy chaining; the reference
    re-executes upstream stages per downstream iterator, SURVEY 3.6 -- that is its
    stock code path and is timed as such)."""
    seed, rows, nchunks, chunk = args
    from oracle import refload

    if refload.load() is None:
        raise RuntimeError("reference not installed")
    from openseize import producer
    from openseize.filtering.fir import Kaiser
    from openseize.filtering.iir import Notch
    from openseize.resampling.resampling import downsample
    from openseize.spectra.estimators import psd

    rng = np.random.default_rng([0, seed])
    pool = [rng.standard_normal((rows, chunk)) for _ in range(2)]

    def source():
        for i in range(nchunks):
            yield pool[i % 2]

    t0 = time.perf_counter()
    pro = producer(source, chunk, -1, shape=(rows, nchunks * chunk))
    p1 = Notch(fstop=60, width=6, fs=FS)(pro, chunk, axis=-1, dephase=True)
    p2 = Kaiser(fpass=500, fstop=600, fs=FS)(p1, chunk, axis=-1, mode="same")
    p3 = downsample(p2, M_DEC, FS, chunk, axis=-1)
    fs2 = FS // M_DEC
    cnt, _, est = psd(p3, fs2, axis=-1, resolution=fs2 / NFFT)
    return time.perf_counter() - t0, float(est.sum()), cnt


class CpuArm:
    """The CPU baseline: every host core runs the pipeline on its own channel block
    (the reference is single-process, single-thread; channel blocks are how it spreads
    over cores).  kind "reference" = the unmodified openseize
    from baseline/_ref; kind "port" = the oracle restatement when that is absent."""

    def __init__(self, cores):
        import multiprocessing as mp

        from oracle import refload

        self.cores = cores
        self.kind = "reference" if refload.location() is not None else "port"
        self.pool = mp.get_context("fork").Pool(cores) if cores > 1 else None

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()

    def step(self, small=False):
        """One bounded sample of the workload on all cores: (rate, seconds, description)."""
        if self.kind == "port":
            n, chunk = (300_000, 100_000) if small else (1_500_000, 250_000)
            jobs = [(i, 2, n, chunk) for i in range(self.cores)]
            fn, total = _cpu_worker, self.cores * 2 * n
            sample = ("%d workers x 2 ch x %d samples (%.0f s of 30 kHz signal), chunksize %d, "
                      "oracle port: each stage runs once" % (self.cores, n, n / FS, chunk))
        else:
            nchunks = 3 if small else REF_CHUNKS
            jobs = [(i, 1, nchunks, CHUNK) for i in range(self.cores)]
            fn, total = _ref_worker, self.cores * nchunks * CHUNK
            sample = ("%d workers x 1 ch x %d chunks of %d samples (%.0f s of 30 kHz signal), "
                      "chunksize 1e6, unmodified openseize %s through its public API"
                      % (self.cores, nchunks, CHUNK, nchunks * CHUNK / FS, _ref_version()))
        t0 = time.perf_counter()
        res = self.pool.map(fn, jobs) if self.pool is not None else [fn(jobs[0])]
        wall = time.perf_counter() - t0
        assert all(np.isfinite(r[1]) for r in res)
        return total / wall, wall, sample


def _ref_version():
    try:
        from importlib import metadata

        from oracle import refload

        dist = [d for d in metadata.distributions(path=[refload.INSTALLED])
                if d.metadata["Name"] == "openseize"]
        return dist[0].version if dist else "(checkout)"
    except Exception:
        return ""


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    arm = CpuArm(cores)
    rates = []
    for _ in range(args.warmup):
        arm.step(small=True)
    t0 = time.perf_counter()
    sample = ""
    for _ in range(args.steps):
        rate, wall, sample = arm.step()
        rates.append(rate)
    total = time.perf_counter() - t0
    arm.close()
    value = float(np.median(rates))
    one_arm = CpuArm(1)
    one, _, one_sample = one_arm.step()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "channel-samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / max(args.steps, 1), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD},
        "cpu_baseline": {"value": value, "unit": "channel-samples/s", "cores": cores,
                         "kind": arm.kind, "sample": sample,
                         "single_core": {"value": one, "unit": "channel-samples/s",
                                         "sample": one_sample}},
        "e2e": {"value": value, "unit": "channel-samples/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def named_kernels(rows, chunk, hbm_peak):
    """The two kernels BASELINE.json's north_star names (Kaiser-FIR oaconvolve
    and Welch PSD), timed alone on an HBM-resident 256 x 1e6 chunk (2 GB, >> L2)
    with CUDA events: 3 warm-up + 5 timed launches each."""
    import scipy.signal as sps
    import torch

    from openseize_b200.core import device as dv
    from openseize_b200.filtering.fir import Kaiser

    gen = torch.Generator(device="cuda").manual_seed(7)
    out = {}

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ms = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        return float(np.mean(ms))

    def entry(ms, ch_samples, bytes_per):
        gbs = ch_samples * bytes_per / (ms * 1e-3) / 1e9
        return {"ms": ms, "channel_samples_per_s": ch_samples / (ms * 1e-3),
                "alg_bytes_per_sample": bytes_per, "achieved_GBps": gbs, "frac": gbs / hbm_peak}

    for fs, label in ((30000, "fir_oaconvolve_kaiser671"), (5000, "fir_oaconvolve_kaiser113")):
        taps = Kaiser(fpass=500, fstop=600, fs=fs).coeffs
        plan = dv.FirPlan(taps)
        x = torch.randn((rows, chunk + len(taps) - 1), dtype=torch.float64, device="cuda",
                        generator=gen)
        y = torch.empty((rows, chunk), dtype=torch.float64, device="cuda")
        out[label] = entry(timed(lambda: plan.run(x, chunk, out=y)), rows * chunk, 16)
        out[label]["taps"] = int(len(taps))
        # opt-in float32 arithmetic (float64 in and out, ~5e-7 of the output peak)
        plan32 = dv.FirPlan(taps, 3)
        out[label + "_f32compute"] = entry(timed(lambda: plan32.run(x, chunk, out=y)),
                                           rows * chunk, 16)
        out[label + "_f32compute"]["taps"] = int(len(taps))
        del x, y
    w = sps.get_window("hann", NFFT)
    plan = dv.SpecPlan(NFFT, NFFT // 2, w, "constant", 1.0 / (FS * float(np.sum(w ** 2))))
    x = torch.randn((rows, chunk), dtype=torch.float64, device="cuda", generator=gen)
    nseg = plan.nseg_available(chunk)
    acc = dv.zeros((rows, NFFT // 2 + 1))
    out["welch_psd_nfft4096"] = entry(timed(lambda: plan.welch_accum(x, nseg, acc)),
                                      rows * nseg * plan.stride, 8)
    # opt-in float32 arithmetic (float64 samples in, float64 sums out)
    plan32 = dv.SpecPlan(NFFT, NFFT // 2, w, "constant", 1.0 / (FS * float(np.sum(w ** 2))),
                         "float32")
    acc.zero_()
    out["welch_psd_nfft4096_f32compute"] = entry(
        timed(lambda: plan32.welch_accum(x, nseg, acc)), rows * nseg * plan.stride, 8)
    del x
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def bind_to_gpu_cpus(index):
    """Pin this rank to the CPU cores next to its GPU (NVML's ideal affinity) before
    it allocates pinned memory: at N = 8 the end-to-end rate is bound by host DRAM
    and the inter-socket link, and pinned chunks should live on the GPU's NUMA node."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


def parity_check(pool, chunk, nchunks=5, nrows=2):
    """The measured pipeline against the CPU oracle on the first `nrows` channels of
    the SAME device pool, `nchunks` chunks of the same cyclic sequence (outside the
    timed region): max PSD bin error relative to each channel's largest bin, and
    the decimated stream's error relative to its peak.  Tolerance 1e-9 (north_star)."""
    import oracle
    import scipy.signal as sps
    from openseize_b200 import producer
    from oracle.chunked import _kaiser_lowpass

    host = [p[:nrows].cpu().numpy() for p in pool]
    x = np.concatenate([host[i % len(host)] for i in range(nchunks)], -1)
    pro = producer(x, chunk, -1)
    dec = build_pipeline(pro, chunk)
    cnt, freqs, est = run_psd(dec)
    got = build_pipeline(producer(x, chunk, -1), chunk).to_array()
    b, a = sps.iirnotch(60, 60 / 6, fs=FS)
    r1 = np.concatenate(oracle.filtfilt(x, (b, a), chunk, -1), -1)
    taps = _kaiser_lowpass(500, 600, FS, 1.0, 40.0)
    r2 = np.concatenate(oracle.oaconvolve(r1, taps, chunk, -1, "same"), -1)
    r3 = np.concatenate(oracle.polyphase_resample(r2, 1, M_DEC, FS, chunk, -1), -1)
    fs2 = FS // M_DEC
    rc, rf, rp = oracle.welch_psd(r3, fs2, -1, fs2 / NFFT)
    assert cnt == rc and np.array_equal(freqs, rf), (cnt, rc)
    psd_err = float(np.max(np.abs(est - rp) / np.max(rp, axis=-1, keepdims=True)))
    dec_err = float(np.max(np.abs(got - r3)) / np.max(np.abs(r3)))
    return {"psd_max_rel_err": psd_err, "decimated_max_rel_err": dec_err, "segments": int(cnt),
            "channels": nrows, "chunks": nchunks, "tolerance": 1e-9,
            "against": "oracle (numpy/scipy restatement pinned to the reference), same pool"}


def h2d_roof(rows, chunk, barrier, max_over_ranks, world, reps=4):
    """What the box delivers for plain pinned host -> device copies with all ranks
    copying at once (one cudaMemcpyAsync per chunk, no kernels): the roof of the
    end-to-end number.  GB/s summed over ranks."""
    import torch

    h = torch.empty((rows, chunk), dtype=torch.float64, pin_memory=True)
    h.numpy()[:] = 1.0
    d = torch.empty((rows, chunk), dtype=torch.float64, device="cuda")
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    secs = max_over_ranks(a.elapsed_time(b) * 1e-3)
    del h, d
    return world * reps * rows * chunk * 8 / secs / 1e9


def run_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cpu_line = None
    if not args.no_cpu and world == 1:
        # before CUDA is initialised: the worker pool forks
        cores = host_cores()
        arm = CpuArm(cores)
        arm.step(small=True)
        rate, wall, sample = arm.step()
        arm.close()
        one_arm = CpuArm(1)
        one, _, one_sample = one_arm.step()
        cpu_line = {"value": rate, "unit": "channel-samples/s", "cores": cores,
                    "kind": arm.kind, "sample": sample, "seconds": wall,
                    # the reference as shipped is single-process, single-thread
                    "single_core": {"value": one, "unit": "channel-samples/s",
                                    "sample": one_sample}}
    import torch

    if world > 1:
        bind_to_gpu_cpus(local)
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    from openseize_b200 import _abi
    from openseize_b200.core import device as dv

    dv.require_cuda()
    # north_star's split: ONE 256-channel recording, channel blocks of 256 / N per GPU
    # (strong scaling, no data-path collective).  --rows overrides the recording's width.
    total_rows, chunk = args.rows, args.chunk
    if total_rows % world:
        raise SystemExit("--rows must be divisible by the number of GPUs")
    rows = total_rows // world
    W, K, CPS = args.warmup, args.steps, args.chunks_per_step
    nchunks = (W + K) * CPS + TAIL
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback 6.65 TB/s"

    def timed_pass(pool, nrows, sampler=None, timers_out=None, launches=None, only=None):
        """W warm-up + K timed steps of CPS chunks each, device-resident pool.  `only`:
        names of the launches bracketed by CUDA events (None: every launch)."""
        marks = Marks(W * CPS, K * CPS, barrier)
        use_timers = timers_out is not None and os.environ.get("OSZ_BENCH_TIMERS", "1") == "1"

        def on_start():
            if sampler:
                sampler.mark_start()
            if launches is not None:
                launches["a"] = _abi.launch_count()
            dv.TIMERS = {} if use_timers else None
            dv.TIMERS_ONLY = set(only) if only else None

        def on_stop():
            if launches is not None:
                launches["b"] = _abi.launch_count()
            if timers_out is not None:
                timers_out.update(dv.TIMERS or {})
            dv.TIMERS = None
            dv.TIMERS_ONLY = None
            if sampler:
                sampler.mark_stop()

        marks.on_start, marks.on_stop = on_start, on_stop
        src = device_source(pool, nrows, chunk, nchunks, marks)
        cnt, freqs, est = run_psd(build_pipeline(src, chunk))
        torch.cuda.synchronize()
        return marks, est

    # ---- value: HBM-resident chunk pool -------------------------------------
    # every rank draws the whole recording's pool with the same seed and keeps its
    # own channel block, so the N-GPU job processes the same 256 channels as N = 1
    gen = torch.Generator(device="cuda").manual_seed(1234)
    pool = []
    for _ in range(2):
        full = torch.randn((total_rows, chunk), dtype=torch.float64, device="cuda", generator=gen)
        pool.append(full[rank * rows:(rank + 1) * rows].clone() if world > 1 else full)
        del full
    torch.cuda.empty_cache()
    sampler = ClockSampler(local, enabled=rank == 0 and os.environ.get("OSZ_BENCH_CLOCKS", "1") == "1",
                           period_ms=45 if world == 1 else 50)
    launches = {}
    sampler.start()
    # Untimed pre-warm, then a barrier: a fresh box pages the CUDA libraries in and
    # ramps its clocks during the first second, and with one rank per GPU the ranks
    # reach steady state at different moments.  The contract's W warm-up steps and K
    # timed steps follow unchanged.
    prewarm = float(os.environ.get("OSZ_BENCH_PREWARM", "0.4" if world > 1 else "0"))
    t_min, t_max = time.perf_counter() + prewarm, time.perf_counter() + (3.0 if prewarm else 0.0)
    while time.perf_counter() < t_min or (sampler.enabled and not sampler.ready
                                          and time.perf_counter() < t_max):
        pre = DeviceSourceOnce(pool, rows, chunk, 4)
        run_psd(build_pipeline(pre.producer(), chunk))
        torch.cuda.synchronize()
    barrier()
    # Two passes of the same W + K steps.  The first brackets EVERY launch with CUDA events
    # (per-kernel table, stage accounting, which kernel dominates); the second is the one
    # `value` comes from and brackets only the dominant kernel's launches -- the events that
    # roofline.achieved is computed from -- because two event records around each of the
    # ~12 launches of a chunk cost 5 % of the step at 32 rows per GPU (1.80 -> 1.70 ms).
    timers_all = {}
    timed_pass(pool, rows, None, timers_all, None)
    dom_name = (max(timers_all, key=lambda k: sum(a.elapsed_time(b) for a, b, _ in timers_all[k]))
                if timers_all else None)
    barrier()
    timers = {}
    marks, est = timed_pass(pool, rows, sampler, timers, launches,
                            only=[dom_name] if dom_name else None)
    sampler.stop()
    secs = max_over_ranks(marks.seconds())
    step_samples = total_rows * chunk * CPS            # whole job, all ranks
    value = K * step_samples / secs
    per_rank = None
    if dist is not None:
        # every rank's own step time and the sum of its kernel times (diagnostic)
        own = torch.tensor([1e3 * marks.own_seconds() / K,
                            sum(a.elapsed_time(b) for recs in timers_all.values()
                                for a, b, _ in recs) / K], dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(own) for _ in range(world)]
        dist.all_gather(allr, own)
        per_rank = {"ms_per_step": [round(float(t[0]), 3) for t in allr],
                    "kernel_ms_per_step": [round(float(t[1]), 3) for t in allr]}
    assert np.all(np.isfinite(est)) and est.shape == (rows, NFFT // 2 + 1)
    # the measured path against the oracle, on channels of the same pool
    parity = parity_check(pool, chunk) if rank == 0 and not args.no_parity else None
    if parity:
        assert parity["psd_max_rel_err"] <= 1e-9 and parity["decimated_max_rel_err"] <= 1e-9, parity

    kernels = {}
    for name, recs in timers_all.items():
        if name in timers:           # the dominant kernel: its events of the timed region
            recs = timers[name]
        ms = [a.elapsed_time(b) for a, b, _ in recs]
        by = [c for _, _, c in recs]
        kernels[name] = {"launches": len(recs), "ms_total": float(np.sum(ms)),
                         "ms_per_step": float(np.sum(ms)) / K,
                         "alg_GBps": float(np.sum(by) / (np.sum(ms) * 1e-3) / 1e9)}
    dom = max(kernels, key=lambda k: kernels[k]["ms_total"]) if kernels else None
    roofline = None
    if dom:
        # dram__bytes_read + dram__bytes_write per launch of this kernel at this shape,
        # from this round's committed `ncu --set full` capture (profiles/r02_traffic.json;
        # the bench never runs under ncu)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
        if os.path.exists(tpath) and rows == ROWS and chunk == CHUNK:
            traffic = json.load(open(tpath)).get(dom, {}).get("bytes_per_launch")
        per_launch = float(np.mean([c for _, _, c in (timers.get(dom) or timers_all[dom])]))
        roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["alg_GBps"],
                    "peak": hbm_peak, "unit": "GB/s",
                    "frac": kernels[dom]["alg_GBps"] / hbm_peak, "traffic": traffic,
                    "alg_bytes_per_launch": per_launch,
                    "peak_source": peak_src,
                    "share_of_step": kernels[dom]["ms_total"] / (secs * 1e3)}
    # stage-level accounting by SURVEY 8(d)'s algorithmic bytes (a forward-backward
    # filter is ONE read and ONE write per sample however many passes it makes)
    stages = None
    if kernels:
        per_step = rows * chunk * CPS

        def stage(names, alg_bytes):
            ms = sum(kernels[n]["ms_per_step"] for n in names if n in kernels)
            if ms <= 0:
                return None
            gbs = alg_bytes * per_step / (ms * 1e-3) / 1e9
            return {"kernels": [n for n in names if n in kernels], "ms_per_step": ms,
                    "alg_bytes_per_sample": alg_bytes, "achieved_GBps": gbs,
                    "frac_of_hbm": gbs / hbm_peak}

        stages = {"notch_filtfilt": stage(["sos", "sos_state", "sos_dec"], 16.0),
                  "fir_decimate": stage(["upfirdn", "fir", "sos_dec"], 8.0 * (1 + 1 / M_DEC)),
                  "welch": None,
                  "pipeline": stage(list(kernels), 8.0 * (1 + 1 / M_DEC))}
    # The same HBM-resident pipeline with the opt-in float32 arithmetic (float64 samples
    # in, float64 results out; the IIR stays float64).  Extra information: the contract's
    # `value` above is the reference's float64.
    f32_mode = None
    if not args.no_f32:
        import openseize_b200

        openseize_b200.set_compute("float32")
        try:
            marks_f, est_f = timed_pass(pool, rows)
        finally:
            openseize_b200.set_compute("float64")
        secs_f = max_over_ranks(marks_f.seconds())
        f32_mode = {"value": K * step_samples / secs_f, "unit": "channel-samples/s",
                    "ms_per_step": 1e3 * secs_f / K,
                    "max_rel_diff_vs_float64": float(np.max(np.abs(est_f - est) / np.max(est, axis=-1,
                                                                                      keepdims=True)))}
    # weak scaling (every rank its own 256-channel recording), as round 1 measured it
    weak = None
    if world > 1 and not args.no_weak:
        gen_w = torch.Generator(device="cuda").manual_seed(99 + rank)
        del pool
        torch.cuda.empty_cache()
        pool_w = [torch.randn((total_rows, chunk), dtype=torch.float64, device="cuda",
                              generator=gen_w) for _ in range(2)]
        marks_w, _ = timed_pass(pool_w, total_rows)
        secs_w = max_over_ranks(marks_w.seconds())
        weak = {"value": world * K * total_rows * chunk * CPS / secs_w,
                "unit": "channel-samples/s", "rows_per_gpu": total_rows,
                "ms_per_step": 1e3 * secs_w / K,
                "note": "every rank streams its own 256-channel recording (N independent replicas)"}
        del pool_w
    else:
        del pool
    torch.cuda.empty_cache()
    named = named_kernels(ROWS, chunk, hbm_peak) if rank == 0 and not args.no_named else None
    barrier()

    # ---- e2e: pinned host chunks through the public API ----------------------
    e2e = None
    if not args.no_e2e:
        roof = h2d_roof(rows, chunk, barrier, max_over_ranks, world)
        host_pool = []
        rng = np.random.default_rng(99 + rank)
        for _ in range(2):
            t = torch.empty((rows, chunk), dtype=torch.float64, pin_memory=True)
            a = t.numpy()
            for r0 in range(0, rows, 32):
                a[r0:r0 + 32] = rng.standard_normal((min(32, rows - r0), chunk))
            host_pool.append(a)
        marks2 = Marks(W * CPS, K * CPS, barrier)
        src2 = host_source(host_pool, rows, chunk, nchunks, marks2)
        cnt2, _, est2 = run_psd(build_pipeline(src2, chunk))
        torch.cuda.synchronize()
        secs2 = max_over_ranks(marks2.seconds())
        wall2 = marks2.wall[(W + K) * CPS] - marks2.wall[W * CPS]
        secs2 = max(secs2, max_over_ranks(wall2))
        e2e_value = K * step_samples / secs2
        e2e = {"value": e2e_value, "unit": "channel-samples/s",
               "h2d_bytes_per_step": rows * chunk * 8 * CPS,
               "d2h_bytes_per_step": int(est2.nbytes * CPS / nchunks),
               "h2d_roof_GBps": roof,
               "frac_of_h2d_roof": e2e_value * 8 / 1e9 / roof,
               "note": "h2d/d2h bytes are per rank; PSD is a streaming reduction: the "
                       "(rows, 2049) result crosses to the host once per recording, its bytes "
                       "are amortised over the steps; h2d_roof = plain pinned cudaMemcpyAsync "
                       "on all ranks at once, summed"}
        # The same pipeline fed float32 and int16 samples (EDF recordings are int16):
        # they cross PCIe in their own width and are widened on the device, so the
        # PCIe roof moves from 8 to 4 and 2 bytes per sample.  Extra information, not
        # the contract's e2e (which stays float64, the dtype of the reference arm).
        if not args.no_narrow and world <= 2:
            e2e["narrow_inputs"] = {}
            for name, tdt, scale in (("float32", torch.float32, 1.0), ("int16", torch.int16, 3000.0)):
                npool = []
                for a in host_pool:
                    t = torch.empty((rows, chunk), dtype=tdt, pin_memory=True)
                    np.multiply(a, scale, out=t.numpy(), casting="unsafe")
                    npool.append(t.numpy())
                marks3 = Marks(W * CPS, K * CPS, barrier)
                src3 = host_source(npool, rows, chunk, nchunks, marks3)
                run_psd(build_pipeline(src3, chunk))
                torch.cuda.synchronize()
                secs3 = max(max_over_ranks(marks3.seconds()),
                            max_over_ranks(marks3.wall[(W + K) * CPS] - marks3.wall[W * CPS]))
                e2e["narrow_inputs"][name] = {
                    "value": K * step_samples / secs3, "unit": "channel-samples/s",
                    "h2d_bytes_per_step": rows * chunk * npool[0].itemsize * CPS}
                del npool
        del host_pool

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "channel-samples/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": 1e3 * secs / K, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "recording_rows": total_rows, "rows_per_gpu": rows,
                       "chunk": chunk, "chunks_per_step": CPS,
                       "step": "%d chunks of %d x %d samples (the whole %d-channel recording; "
                               "each GPU filters its %d-channel block)" % (CPS, total_rows, chunk,
                                                                            total_rows, rows),
                       "sharding": "channel blocks of one recording, one process per GPU, "
                                   "no data-path collective",
                       "l2": "each chunk a rank reads is %.0f MB (> 126 MB L2); pool of 2 chunks"
                             % (rows * chunk * 8 / 1e6)},
            "gpu_launches": int(launches.get("b", 0) - launches.get("a", 0)),
            "clocks": sampler.summary(), "kernels": kernels,
            "kernels_note": "per-kernel CUDA-event times: the dominant kernel (%s) over the timed "
                            "region itself, the others from a pass of the same W + K steps run "
                            "just before it with every launch bracketed" % dom,
            "host_enqueue": marks.host_pace(),
            # sum of our kernels' CUDA-event times per step: with the host a whole timed
            # region ahead, ms_per_step minus this is idle time on the device side
            "kernel_sum_ms_per_step": sum(v["ms_total"] for v in kernels.values()) / K,
        }
        if parity:
            line["parity"] = parity
            line["parity_err"] = parity["psd_max_rel_err"]
        if per_rank:
            line["per_rank"] = per_rank
        if roofline:
            line["roofline"] = roofline
        if stages:
            line["stages"] = stages
        if f32_mode:
            line["float32_compute"] = f32_mode
        if weak:
            line["weak_scaling"] = weak
        if named:
            line["named_kernels"] = named
            # These float64 kernels are bound by the FP64 pipe, not by HBM (ncu: FP64 pipe
            # 67 % / 67 % / 59 % busy with DRAM at 48 / 42 / 20 %, profiles/r02_ncu_summary.md
            # A and r01 C): next to the HBM fraction north_star asks for, the fraction of the
            # FP64 issue peak (SMs x 64 lanes x SM clock) the measured rate corresponds to,
            # from the kernels' FP64 instructions per sample (ncu instruction counts).
            fp64_per_sample = {"fir_oaconvolve_kaiser113": 48.6, "fir_oaconvolve_kaiser671": 56.4,
                               "welch_psd_nfft4096": 54.4}
            clk = (sampler.summary() or {}).get("sm_mhz") or 1965.0
            fp64_peak = torch.cuda.get_device_properties(local).multi_processor_count * 64 * clk * 1e6
            line["roofline_named"] = {}
            for k, v in named.items():
                if "f32" in k:
                    continue
                ent = {"bound": "fp64" if k in fp64_per_sample else "hbm", "frac": v["frac"],
                       "achieved": v["achieved_GBps"], "peak": hbm_peak, "unit": "GB/s"}
                if k in fp64_per_sample:
                    rate = v["channel_samples_per_s"] * fp64_per_sample[k]
                    ent["fp64"] = {"instr_per_sample": fp64_per_sample[k],
                                   "achieved_instr_per_s": rate, "peak_instr_per_s": fp64_peak,
                                   "frac": rate / fp64_peak}
                    ent["frac_note"] = "frac = algorithmic bytes / time / HBM peak (north_star's metric); the kernel's own bound is fp64.frac"
                line["roofline_named"][k] = ent
        if e2e:
            line["e2e"] = e2e
        if cpu_line:
            line["cpu_baseline"] = cpu_line
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.shard == "time":
        run_time_shard(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
