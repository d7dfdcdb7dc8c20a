import torch, time
n = 512*1024*1024
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device='cuda'); d2 = torch.empty(n, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(f, reps=3):
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/reps
def h2d():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
def both(): h2d(); d2h()
h2d(); d2h(); torch.cuda.synchronize()
a=t(h2d); b=t(d2h); c=t(both)
print("H2D %.1f GB/s  D2H %.1f GB/s  both: %.1f GB/s each (%.1f total)"%(n/a/1e9, n/b/1e9, n/c/1e9, 2*n/c/1e9))
