"""Pinned host -> device copy bandwidth of this box (the ceiling of bench.py's e2e number: 92.9 MB of inputs per pair)."""
import torch
for mb in (15, 77, 372):
    h = torch.empty(mb * 1000 * 1000, dtype=torch.uint8).pin_memory()
    d = torch.empty_like(h, device="cuda")
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    print(f"H2D pinned {mb} MB: {mb * 10 / e0.elapsed_time(e1):.1f} GB/s")
