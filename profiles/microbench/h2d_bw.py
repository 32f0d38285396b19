"""Host->device / device->host copy bandwidth from pinned memory on this box (sets the ceiling of bench.py's e2e number)."""
import torch
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, (dst, src) in {"h2d": (d, h), "d2h": (h, d)}.items():
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    print(name, "%.1f GB/s" % (5 * n / e0.elapsed_time(e1) / 1e6))
