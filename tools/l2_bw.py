"""L2 read bandwidth of the GPU (denominator for the attention sweeps, which read L2-resident fp16 memories): torch.sum /
torch.mul_ over buffers that fit the 126 MB L2, timed with CUDA events after a warm-up pass that brings them in."""
import torch
for mb in (8, 16, 32, 64, 96, 256, 1024):
    n = mb * 1024 * 1024 // 4
    x = torch.randn(n, device='cuda')
    for _ in range(3):
        x.sum()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        x.sum()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / reps
    y = x.half()
    for _ in range(3):
        y.sum(dtype=torch.float32)
    e0.record()
    for _ in range(reps):
        y.sum(dtype=torch.float32)
    e1.record()
    torch.cuda.synchronize()
    t2 = e0.elapsed_time(e1) / reps
    print(f'{mb:5d} MB fp32: sum {mb / 1024 / (t * 1e-3):8.1f} GB/s ({t * 1e3:7.1f} us)   {mb // 2:5d} MB fp16: sum '
          f'{mb / 2 / 1024 / (t2 * 1e-3):8.1f} GB/s ({t2 * 1e3:7.1f} us)')
