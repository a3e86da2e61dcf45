"""Development: write-only and copy bandwidth of the GPU (context for the planner kernel's HBM fraction)."""
import torch
x = torch.empty(1 << 28, dtype=torch.float32, device="cuda")     # 1 GiB
y = torch.empty_like(x)
def t(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
ms = t(lambda: x.zero_()); print(f"memset 1 GiB: {ms:.3f} ms -> {x.numel() * 4 / ms / 1e6:.0f} GB/s written")
ms = t(lambda: x.fill_(1.5)); print(f"fill 1 GiB: {ms:.3f} ms -> {x.numel() * 4 / ms / 1e6:.0f} GB/s written")
ms = t(lambda: y.copy_(x)); print(f"copy 1 GiB: {ms:.3f} ms -> {2 * x.numel() * 4 / ms / 1e6:.0f} GB/s read+written")
z = torch.empty(134_600_000 // 4, dtype=torch.float32, device="cuda")
ms = t(lambda: z.fill_(1.5), 50); print(f"fill 134.6 MB: {ms*1e3:.1f} us -> {z.numel() * 4 / ms / 1e6:.0f} GB/s written")
