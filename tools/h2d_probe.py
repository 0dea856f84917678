"""Raw pinned host->device bandwidth on this box (development aid for the e2e number)."""
import torch
for mb in (64, 256, 1232):
    h = torch.empty(mb * 1024 * 1024, dtype=torch.uint8).pin_memory()
    d = torch.empty_like(h, device="cuda")
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    print("H2D %5d MiB: %.1f GB/s" % (mb, 5 * h.numel() / e0.elapsed_time(e1) / 1e6))
    h2 = torch.empty_like(h).pin_memory()
    e0.record()
    for _ in range(5):
        h2.copy_(d, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    print("D2H %5d MiB: %.1f GB/s" % (mb, 5 * h.numel() / e0.elapsed_time(e1) / 1e6))
