"""Runs the three largest gate products of the bench workload a few times (for ncu --set full -k regex:gemm_tc)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avsr_tf1_b200 import ops
for (ta, tb, M, N, K) in [(0, 0, 76800, 1024, 256), (0, 1, 76800, 256, 1024), (1, 0, 256, 1024, 76800), (0, 0, 19200, 1024, 3888)]:
    a = ops.round_tf32(torch.randn((K, M) if ta else (M, K), device='cuda'))
    b = ops.round_tf32(torch.randn((N, K) if tb else (K, N), device='cuda'))
    c = torch.empty(M, N, device='cuda')
    for _ in range(3):
        ops.gemm(a, b, c, ta=bool(ta), tb=bool(tb))
    torch.cuda.synchronize()
