"""Event-timed attention-LSTM layer (persistent kernels) at the bench shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avsr_tf1_b200 import ops
B, H = 256, 256


class Drop:  # what ops.RnnSeq reads of a layers.DropState (keep 0.9 / 0.9 / 0.9: the reference default)
    def __init__(self):
        self.rng = torch.tensor([1234, 5], dtype=torch.int32, device='cuda')
        self.stream = 8
        self.thr_in = self.thr_state = self.thr_out = ops.keep_threshold(0.9)


ops.kernel_timing(True)
CASES = [('cross-modal', 300, 256, 75, None), ('decoder', 41, 128, 300, None),
         ('cross-modal+dropout', 300, 256, 75, Drop()), ('decoder+dropout', 41, 128, 300, Drop())]
if len(sys.argv) > 1 and sys.argv[1] == 'sweep':  # memory-length sweep: what of a step is the memory sweeps, what the chain
    CASES = [('xmodal+dropout Tm=%d' % tm, 300, 256, tm, Drop()) for tm in (4, 20, 40, 75, 150, 300)]
    CASES += [('xmodal Tm=%d' % tm, 300, 256, tm, None) for tm in (4, 75)]
for (name, T, Dx, Tm, drop) in CASES:
    A = Dm = 256
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
    best = [1e9, 1e9]
    for rep in range(3):
        gates = gates0.clone()
        mb = ops.MechBuffers('scaled_luong', values, keys, mlen, Wl, g=g)
        rnn = ops.RnnSeq(T, B, H, lens, gates, W[Dx:], [mb], True, drop=drop)
        mb.dkeys, mb.dvalues = torch.zeros_like(keys), torch.zeros_like(values)
        mb.dWl, mb.dg = torch.zeros_like(Wl), torch.zeros(1, device='cuda')
        gW = torch.zeros_like(W)
        dout = torch.randn(T, B, A, device='cuda') * 1e-3
        rnn.grad_scale = 1024.0
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(); rnn.forward(); e1.record(); rnn.backward(dout, gW[Dx:]); e2.record()
        torch.cuda.synchronize()
        best = [min(best[0], e0.elapsed_time(e1) * 1e3), min(best[1], e1.elapsed_time(e2) * 1e3)]
    kt = ops.kernel_times()
    ops.kernel_timing(True)
    print(f'{name:20s} kernels only (3 reps): fwd {kt["attn_lstm_fwd"][0] / 3 * 1e3:8.1f} us  bwd {kt["attn_lstm_bwd"][0] / 3 * 1e3:8.1f} us')
    print(f'{name:20s} T={T:3d} Tm={Tm:3d}: fwd {best[0]:8.1f} us ({best[0] / T:6.2f}/step)  bwd {best[1]:8.1f} us ({best[1] / T:6.2f}/step)', flush=True)
