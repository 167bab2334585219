"""Quick tcgen05-GEMM check against fp64 for the four operand layouts (prints, never asserts)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from avsr_tf1_b200 import ops

def run(M, N, K, ta, tb, beta, bias, tc):
    ops.set_tensor_cores(tc)
    g = torch.Generator(device='cuda').manual_seed(M + N + K)
    A = torch.randn((K, M) if ta else (M, K), device='cuda', generator=g)
    B = torch.randn((N, K) if tb else (K, N), device='cuda', generator=g)
    C0 = torch.randn(M, N, device='cuda', generator=g)
    bv = torch.randn(N, device='cuda', generator=g) if bias else None
    ref = (A.double().t() if ta else A.double()) @ (B.double().t() if tb else B.double())
    if beta: ref = ref + C0.double()
    if bias: ref = ref + bv.double()
    C = C0.clone()
    ops.gemm(A, B, C, ta=ta, tb=tb, beta=beta, bias=bv)
    torch.cuda.synchronize()
    err = (C.double() - ref).abs().max().item() / ref.abs().max().item()
    return err

cases = [(19200, 1024, 256), (1000, 80, 1024), (256, 1024, 19200),
         (80, 1024, 7680), (3888, 1024, 1920), (1920, 3888, 1024), (256, 1024, 512), (300, 200, 100)]
for (M, N, K) in cases:
    for ta in (False, True):
        for tb in (False, True):
            for beta, bias in ((0.0, False), (1.0, False), (0.0, True)):
                e_tc = run(M, N, K, ta, tb, beta, bias, True)
                e_f = run(M, N, K, ta, tb, beta, bias, False)
                flag = 'OK ' if e_tc < 2e-3 else 'BAD'
                print(f'{flag} M={M} N={N} K={K} ta={int(ta)} tb={int(tb)} beta={beta} bias={int(bias)}  '
                      f'tf32 err={e_tc:.2e}  fp32 err={e_f:.2e}', flush=True)
