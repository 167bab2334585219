"""Pinned host -> device copy bandwidth of the box (context for bench.py's e2e number)."""
import torch
for mb in (32, 298):
    h = torch.empty(mb * 1024 * 1024 // 4, dtype=torch.float32).pin_memory()
    d = torch.empty_like(h, device='cuda')
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f'H2D {mb} MiB pinned: {ms:.2f} ms  {h.numel() * 4 / ms / 1e6:.1f} GB/s', flush=True)
