import torch, time
for mb in (1, 6, 23, 64):
    n = mb * 1024 * 1024
    h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for name, a, b in (("H2D", d, h), ("D2H", h, d)):
        for _ in range(3): a.copy_(b, non_blocking=True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10): a.copy_(b, non_blocking=True)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
        print(f"{name} {mb} MiB: {dt*1e3:.3f} ms  {n/dt/1e9:.1f} GB/s")
