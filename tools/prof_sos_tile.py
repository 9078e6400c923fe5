"""Scratch: a few launches of the tiled notch scan for ncu."""
import os, sys
import numpy as np, scipy.signal as sps, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openseize_b200.core import device as dv
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 256
b, a = sps.iirnotch(60, 10, fs=30000)
plan = dv.SosPlan(np.concatenate([b, a])[None])
x = torch.randn((rows, 1_000_000), dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
st = dv.zeros((rows, 1, 2))
for _ in range(3):
    plan.run(x, st, out=y)
torch.cuda.synchronize()
