"""Per-phase clocks of the persistent LSTM forward kernel (AVSR_LP_DEBUG=1)."""
import os, sys
os.environ['AVSR_LP_DEBUG'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avsr_tf1_b200 import ops
for (T, B, H, I) in [(300, 256, 256, 256), (75, 256, 256, 256), (300, 64, 256, 80)]:
    x = ops.round_tf32(torch.randn(T, B, I, device='cuda'))
    W = ops.round_tf32(torch.randn(I + H, 4 * H, device='cuda') / (I + H) ** 0.5)
    lens = torch.full((B,), T, dtype=torch.int32, device='cuda')
    gates = torch.empty(T, B, 4 * H, device='cuda')
    ops.gemm(x.view(T * B, I), W[:I], gates.view(T * B, 4 * H))
    rnn = ops.RnnSeq(T, B, H, lens, gates, W[I:])
    for _ in range(2):
        rnn.forward()
    torch.cuda.synchronize()
