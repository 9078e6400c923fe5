"""Time one forward SOS pass over 1e6 samples for 4...128 rows (few-channel recordings):
the launch policy of osz_sos_exec_f64 cuts rows into concurrent spans (warm-up or exact
two-pass split).  OSZ_SOS_SPLIT / OSZ_SOS_EXACT force a span count.  Usage on a GPU box:
    python tools/sos_split_sweep.py butter8|notch"""
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, scipy.signal as sps, torch
from openseize_b200.core import device as dv
dv.require_cuda()
kind = sys.argv[1]
if kind=="butter8": sos = sps.butter(8, [1, 100], btype="bandpass", fs=5000, output="sos")
else:
    b,a = sps.iirnotch(60,10,fs=30000); sos=np.concatenate([b,a])[None]
plan = dv.SosPlan(sos)
n=1_000_000
for rows in (4,16,32,64,128):
    x=torch.randn((rows,n),dtype=torch.float64,device="cuda"); y=torch.empty_like(x); st=dv.zeros((rows,sos.shape[0],2))
    for _ in range(2): plan.run(x,st,out=y)
    torch.cuda.synchronize(); ts=[]
    for _ in range(5):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); plan.run(x,st,out=y); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    print("%s rows=%3d SPLIT=%s EXACT=%s  %.3f ms  %.1f Gsamp/s"%(kind,rows,os.environ.get("OSZ_SOS_SPLIT"),os.environ.get("OSZ_SOS_EXACT"),np.median(ts),rows*n/np.median(ts)/1e6))
