"""Experiment: can the forward output F of a notch filtfilt stay in the
126 MB L2 between the forward and the backward pass?  Row groups of G channels are run
forward then backward through ONE reused G-row F buffer, against the whole-chunk schedule
(forward over all 256 rows -> 2 GB of F through HBM -> backward).  Captured in a CUDA graph
so the host launch pace does not limit the grouped schedule.

    python tools/l2_filtfilt.py            # prints one line per schedule
"""
import os
import sys

import numpy as np
import scipy.signal as sps
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openseize_b200.core import device as dv  # noqa: E402

rows, n = 256, 1_000_000
b, a = sps.iirnotch(60, 10, fs=30000)
plan = dv.SosPlan(np.concatenate([b, a])[None])
x = torch.randn((rows, n), dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
F = torch.empty_like(x)
st_f = dv.zeros((rows, 1, 2))
st_b = dv.zeros((rows, 1, 2))


def whole():
    plan.run(x, st_f, out=F)
    plan.run(F, st_b, reverse=True, out=y)


def grouped(G):
    Fs = F[:G]

    def go():
        for g in range(0, rows, G):
            plan.run(x[g:g + G], st_f[g:g + G], out=Fs)
            plan.run(Fs, st_b[g:g + G], reverse=True, out=y[g:g + G])
    return go


def timed(fn, label, reps=5):
    fn()
    torch.cuda.synchronize()
    mode = "eager"
    run = fn
    try:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fn()
        torch.cuda.current_stream().wait_stream(s)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            fn()
        run = graph.replay
        mode = "graph"
    except Exception as exc:  # capture not possible: time the eager launches
        print("  (graph capture failed for %s: %s)" % (label, str(exc)[:120]))
        torch.cuda.synchronize()
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = float(np.median(ts))
    print("L2_FILTFILT %-28s %-5s %7.3f ms  %6.1f G samples/s  %5.1f %% of the 16 B/sample roofline"
          % (label, mode, t, rows * n / t / 1e6, 100 * rows * n * 16 / (t * 1e-3) / 6534.1e9), flush=True)


if __name__ == "__main__":
    which = [a for a in sys.argv[1:] if a != "--once"] or ["whole", "64", "32", "16", "8"]
    if "--once" in sys.argv:        # one eager pass, for ncu --cache-control none
        for w in which:
            (whole if w == "whole" else grouped(int(w)))()
        torch.cuda.synchronize()
        sys.exit(0)
    for w in which:
        if w == "whole":
            timed(whole, "whole chunk (256 rows)")
        else:
            timed(grouped(int(w)), "groups of %d rows" % int(w))
