"""One forward + backward of (1) the dual-attention (WLAS) decoder layer at the config-4 shape and (2) the fused resnet_cnn
front-end at 64 x 75 crops inside a cudaProfilerStart/Stop range - the target of `ncu --set full -k regex:wlas|conv_mma`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avsr_tf1_b200 import ops
from avsr_tf1_b200.layers import BuildContext
from avsr_tf1_b200.params import ParamStore
from avsr_tf1_b200.video import ResNetCNN

B, H, T, Dx, A = 128, 256, 41, 128, 256
TMS = (75, 300)


class Drop:
    def __init__(self, stream):
        self.rng = torch.tensor([1234, 5], dtype=torch.int32, device='cuda')
        self.stream = stream
        self.thr_in = self.thr_state = self.thr_out = ops.keep_threshold(0.9)


x = ops.round_tf32(torch.randn(T, B, Dx, device='cuda'))
W = ops.round_tf32(torch.randn(Dx + 2 * A + H, 4 * H, device='cuda') / (Dx + 2 * A + H) ** 0.5)
lens = torch.full((B,), T, dtype=torch.int32, device='cuda')
gates0 = torch.empty(T, B, 4 * H, device='cuda')
ops.gemm(x.view(T * B, Dx), W[:Dx], gates0.view(T * B, 4 * H))
mem = []
for Tm in TMS:
    Wl = ops.round_tf32(torch.randn(H + 256, A, device='cuda') / (H + 256) ** 0.5)
    Wm = ops.round_tf32(torch.randn(256, A, device='cuda') / 16.0)
    values = ops.round_tf32(torch.tanh(torch.randn(Tm, B, 256, device='cuda')))
    keys = torch.empty(Tm, B, A, device='cuda')
    ops.gemm(values.view(Tm * B, 256), Wm, keys.view(Tm * B, A))
    mem.append((values, keys, torch.full((B,), Tm, dtype=torch.int32, device='cuda'), Wl))

ctx = BuildContext()
cnn = ResNetCNN(ctx, 36, 36, 3)
ctx.store = ParamStore(ctx.specs, device='cuda', with_optimizer=True)
ctx.store.initialize(7)
N = 64 * 75
frames = torch.rand(N, 36, 36, 3, device='cuda') * 2 - 1
dfeat = torch.randn(N, 128, device='cuda') * 1e-3


def run():
    bufs = []
    for values, keys, mlen, Wl in mem:
        mb = ops.MechBuffers('scaled_luong', values, keys, mlen, Wl, g=torch.ones(1, device='cuda'))
        mb.dkeys, mb.dvalues = torch.zeros_like(keys), torch.zeros_like(values)
        mb.dWl, mb.dg = torch.zeros_like(Wl), torch.zeros(1, device='cuda')
        bufs.append(mb)
    rnn = ops.RnnSeq(T, B, H, lens, gates0.clone(), W[Dx:], bufs, True, drop=Drop(8))
    rnn.grad_scale = 1024.0
    rnn.forward()
    rnn.backward(torch.randn(T, B, 2 * A, device='cuda') * 1e-3, torch.zeros_like(W)[Dx:])
    cnn.forward(frames, True)
    cnn.backward(dfeat)
    torch.cuda.synchronize()


run()
torch.cuda.profiler.start()
run()
torch.cuda.profiler.stop()
