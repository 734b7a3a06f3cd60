"""Pinned host->device bandwidth for the bench's image payload (debug aid)."""
import torch, time
n = 32 * 752 * 480
h = [torch.zeros(n, dtype=torch.uint8).pin_memory() for _ in range(2)]
d = [torch.zeros(n, dtype=torch.uint8, device="cuda") for _ in range(2)]
s = torch.cuda.Stream()
for rep in range(3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        e0.record(s)
        for i in range(20):
            d[0].copy_(h[0], non_blocking=True); d[1].copy_(h[1], non_blocking=True)
        e1.record(s)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"H2D pinned: {40 * n / ms / 1e6:.1f} GB/s ({ms / 20:.3f} ms per 2 x {n / 1e6:.1f} MB)")
