import torch, time
n = 78796800
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device='cuda')
o = torch.empty(16345088, dtype=torch.uint8).pin_memory()
od = torch.empty(16345088, dtype=torch.uint8, device='cuda')
s = torch.cuda.Stream(); s2 = torch.cuda.Stream()
for _ in range(3): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(10): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t) / 10
print("H2D %.2f ms  %.1f GB/s" % (dt * 1e3, n / dt / 1e9))
t = time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): o.copy_(od, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t) / 10
print("H2D+D2H duplex %.2f ms" % (dt * 1e3))
