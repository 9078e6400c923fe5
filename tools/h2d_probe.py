import torch, time
n = 256 * 1_000_000
h = torch.empty(n, dtype=torch.float64, pin_memory=True)
d = torch.empty(n, dtype=torch.float64, device="cuda")
for _ in range(2):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); d.copy_(h, non_blocking=True); b.record(); torch.cuda.synchronize()
print("H2D pinned 2.048 GB: %.1f ms -> %.1f GB/s" % (a.elapsed_time(b), n * 8 / a.elapsed_time(b) / 1e6))
a.record(); h.copy_(d, non_blocking=True); b.record(); torch.cuda.synchronize()
print("D2H pinned 2.048 GB: %.1f ms -> %.1f GB/s" % (a.elapsed_time(b), n * 8 / a.elapsed_time(b) / 1e6))
