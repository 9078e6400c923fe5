import torch, numpy as np
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts=[]
    for _ in range(reps):
        a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b)*1e-3)
    return float(np.median(ts))
n = 1<<28   # 2 GiB of float64
x = torch.empty(n, dtype=torch.float64, device="cuda"); y = torch.empty(n, dtype=torch.float64, device="cuda")
t = timeit(lambda: x.zero_()); print("write only  %.0f GB/s" % (n*8/t/1e9))
t = timeit(lambda: y.copy_(x)); print("copy        %.0f GB/s (read+write)" % (2*n*8/t/1e9))
t = timeit(lambda: x.sum()); print("read only   %.0f GB/s" % (n*8/t/1e9))
# 1 read : 2 writes
z = torch.empty(2*n//2, dtype=torch.float64, device="cuda")
h = x[:n//2]
def rw():
    z[:n//2].copy_(h); z[n//2:].copy_(h)
t = timeit(rw); print("2 copies of half: %.0f GB/s" % (2*n*8/t/1e9))
