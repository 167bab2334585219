"""One forward + backward of the cross-modal attention-LSTM layer and of a plain LSTM layer at the bench shapes,
inside a cudaProfilerStart/Stop range - the target of `ncu --set full -k regex:persist4` (tools/gpu_ncu_full.sh)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avsr_tf1_b200 import ops

B, H, T, Dx, Tm, A, Dm = 256, 256, 300, 256, 75, 256, 256
DROP = '--drop' in sys.argv  # the DropoutWrapper kernels (attn_persist4d.cu, lstm_persist4 DROP) instead of the folded ones


class Drop:  # what ops.RnnSeq reads of a layers.DropState (keep 0.9 / 0.9 / 0.9: the reference default)
    def __init__(self, stream):
        self.rng = torch.tensor([1234, 5], dtype=torch.int32, device='cuda')
        self.stream = stream
        self.thr_in = self.thr_state = self.thr_out = ops.keep_threshold(0.9)


x = ops.round_tf32(torch.randn(T, B, Dx, device='cuda'))
W = ops.round_tf32(torch.randn(Dx + A + H, 4 * H, device='cuda') / (Dx + A + H) ** 0.5)
Wl = ops.round_tf32(torch.randn(H + Dm, A, device='cuda') / (H + Dm) ** 0.5)
Wm = ops.round_tf32(torch.randn(Dm, A, device='cuda') / Dm ** 0.5)
g = torch.ones(1, device='cuda')
lens = torch.full((B,), T, dtype=torch.int32, device='cuda')
mlen = torch.full((B,), Tm, dtype=torch.int32, device='cuda')
values = ops.round_tf32(torch.tanh(torch.randn(Tm, B, Dm, device='cuda')))
keys = torch.empty(Tm, B, A, device='cuda')
ops.gemm(values.view(Tm * B, Dm), Wm, keys.view(Tm * B, A))
gates0 = torch.empty(T, B, 4 * H, device='cuda')
ops.gemm(x.view(T * B, Dx), W[:Dx], gates0.view(T * B, 4 * H))
Wp = ops.round_tf32(torch.randn(Dx + H, 4 * H, device='cuda') / (Dx + H) ** 0.5)
gatesp0 = torch.empty(T, B, 4 * H, device='cuda')
ops.gemm(x.view(T * B, Dx), Wp[:Dx], gatesp0.view(T * B, 4 * H))


def run():
    mb = ops.MechBuffers('scaled_luong', values, keys, mlen, Wl, g=g)
    rnn = ops.RnnSeq(T, B, H, lens, gates0.clone(), W[Dx:], [mb], True, drop=Drop(8) if DROP else None)
    mb.dkeys, mb.dvalues = torch.zeros_like(keys), torch.zeros_like(values)
    mb.dWl, mb.dg = torch.zeros_like(Wl), torch.zeros(1, device='cuda')
    gW = torch.zeros_like(W)
    rnn.grad_scale = 1024.0
    rnn.forward()
    rnn.backward(torch.randn(T, B, A, device='cuda') * 1e-3, gW[Dx:])
    plain = ops.RnnSeq(T, B, H, lens, gatesp0.clone(), Wp[Dx:], drop=Drop(12) if DROP else None)
    plain.grad_scale = 1024.0
    plain.forward()
    plain.backward(torch.randn(T, B, H, device='cuda') * 1e-3, torch.zeros(H, 4 * H, device='cuda'))
    torch.cuda.synchronize()


run()
torch.cuda.profiler.start()
run()
torch.cuda.profiler.stop()
