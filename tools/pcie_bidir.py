import torch, time
n_in, n_out = 1220 << 20, 2070 << 20
hi = torch.empty(n_in, dtype=torch.uint8).pin_memory(); ho = torch.empty(n_out, dtype=torch.uint8).pin_memory()
di = torch.empty(n_in, dtype=torch.uint8, device="cuda"); do = torch.empty(n_out, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3
for _ in range(2):
    a = t(lambda: di.copy_(hi, non_blocking=True)); b = t(lambda: ho.copy_(do, non_blocking=True))
    def both():
        with torch.cuda.stream(s1): di.copy_(hi, non_blocking=True)
        with torch.cuda.stream(s2): ho.copy_(do, non_blocking=True)
    c = t(both)
    print(f"H2D {a:.1f} ms ({n_in/a/1e6:.1f} GB/s)  D2H {b:.1f} ms ({n_out/b/1e6:.1f} GB/s)  both {c:.1f} ms ({(n_in+n_out)/c/1e6:.1f} GB/s)")
