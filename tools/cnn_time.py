"""Forward + backward time of the resnet_cnn front-end alone (im2col + dense products: the functional version)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avsr_tf1_b200.layers import BuildContext
from avsr_tf1_b200.params import ParamStore
from avsr_tf1_b200.video import ResNetCNN

ctx = BuildContext()
cnn = ResNetCNN(ctx, 36, 36, 3)
ctx.store = ParamStore(ctx.specs, device='cuda', with_optimizer=True)
ctx.store.initialize(7)
for B in (16, 64, 256):
    N = B * 75
    frames = torch.rand(N, 36, 36, 3, device='cuda') * 2 - 1
    d = torch.randn(N, 128, device='cuda') * 1e-3
    for _ in range(2):
        f = cnn.forward(frames, True); cnn.backward(d)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(); f = cnn.forward(frames, True); e[1].record(); cnn.backward(d); e[2].record()
    torch.cuda.synchronize()
    fw, bw = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
    flop = 11.47e6 * N
    print(f'B={B:4d} ({N} frames): fwd {fw:8.2f} ms  bwd {bw:8.2f} ms  -> {B / ((fw + bw) / 1e3):8.1f} utt/s, '
          f'fwd {flop / fw / 1e9:.2f} TFLOP/s, peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB', flush=True)
    del frames, d, f
    torch.cuda.empty_cache()
