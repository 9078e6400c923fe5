"""Stress test of the tiled single-section scan: many launches of random shapes on TWO
streams at once (so that grids are only partly resident while another kernel holds SMs),
both ways of dealing tiles, checked against the one-CTA-per-row kernel on the same data.

    timeout 300 python tools/stress_sos_tile.py [seconds]
"""
import os
import sys
import time

import numpy as np
import scipy.signal as sps
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openseize_b200.core import device as dv  # noqa: E402


def main(seconds):
    rng = np.random.default_rng(5)
    b, a = sps.iirnotch(60, 10, fs=30000)
    plan = dv.SosPlan(np.concatenate([b, a])[None])
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    t_end = time.time() + seconds
    launches = worst = 0
    while time.time() < t_end:
        jobs = []
        for s in streams:
            rows = int(rng.choice([1, 3, 8, 32, 64, 150, 256, 300]))
            n = int(rng.integers(2, 60)) * 4096 + int(rng.choice([0, 2, 16, 576]))
            reverse = bool(rng.integers(0, 2))
            os.environ["OSZ_SOS_TILE_DEAL"] = str(int(rng.integers(0, 2)))
            with torch.cuda.stream(s):
                x = torch.randn((rows, n), dtype=torch.float64, device="cuda")
                st = torch.zeros((rows, 1, 2), dtype=torch.float64, device="cuda")
                os.environ["OSZ_SOS_TILE"] = "1"
                y = plan.run(x, st, reverse=reverse)
                look = plan.lookahead(np.array([[0.3, -0.2]]), x[:, :min(n, 70000)], reverse=True)
            jobs.append((s, x, y, st, look, reverse))
            launches += 2
        for s, x, y, st, look, reverse in jobs:
            s.synchronize()
            os.environ["OSZ_SOS_TILE"] = "0"
            st2 = torch.zeros_like(st)
            ref = plan.run(x, st2, reverse=reverse)
            torch.cuda.synchronize()
            err = float((y - ref).abs().max() / ref.abs().max())
            serr = float((st - st2).abs().max() / (st2.abs().max() + 1e-300))
            worst = max(worst, err)
            assert err < 1e-10 and serr < 1e-7, (err, serr, tuple(x.shape), reverse)
            assert bool(torch.isfinite(look).all())
    os.environ.pop("OSZ_SOS_TILE", None)
    os.environ.pop("OSZ_SOS_TILE_DEAL", None)
    print("STRESS OK: %d launches on two streams, worst relative error %.2e" % (launches, worst))


if __name__ == "__main__":
    main(float(sys.argv[1]) if len(sys.argv) > 1 else 60.0)
