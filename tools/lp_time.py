"""Event-timed persistent LSTM kernels for several T: separates fixed launch cost from per-step cost."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avsr_tf1_b200 import ops
B, H, I = 256, 256, 256
for T in (1, 8, 32, 75, 150, 300):
    x = ops.round_tf32(torch.randn(T, B, I, device='cuda'))
    W = ops.round_tf32(torch.randn(I + H, 4 * H, device='cuda') / (I + H) ** 0.5)
    lens = torch.full((B,), T, dtype=torch.int32, device='cuda')
    gates0 = torch.empty(T, B, 4 * H, device='cuda')
    ops.gemm(x.view(T * B, I), W[:I], gates0.view(T * B, 4 * H))
    res = []
    for rep in range(4):
        gates = gates0.clone()
        rnn = ops.RnnSeq(T, B, H, lens, gates, W[I:])
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(); rnn.forward(); e1.record()
        gW = torch.zeros(H, 4 * H, device='cuda')
        dout = torch.randn(T, B, H, device='cuda')
        e1.record(); rnn.backward(dout, gW); e2.record()
        torch.cuda.synchronize()
        res.append((e0.elapsed_time(e1) * 1e3, e1.elapsed_time(e2) * 1e3))
    f, b = min(r[0] for r in res), min(r[1] for r in res)
    print(f'T={T:4d}  fwd {f:8.1f} us ({f / T:6.2f}/step)   bwd+dW {b:8.1f} us ({b / T:6.2f}/step)', flush=True)
