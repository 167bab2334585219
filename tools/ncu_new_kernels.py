"""Kernels added late in round 1 at the bench shapes, inside a cudaProfilerStart/Stop range (tools/gpu_ncu_new.sh):
vectorised input-BN apply on the 3888-wide lip-crop features, the tensor-core outer products of the attention backward,
the dropout pass, and the cluster-of-4 LSTM kernels with in-kernel dropout masks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avsr_tf1_b200 import ops

B, H, T, Tv, F = 256, 256, 300, 75, 3888
x = torch.randn(B, Tv, F, device='cuda')
sums = torch.zeros(2 * F, device='cuda')
ops.bn_stats(x.view(B * Tv, F), sums)
gamma, beta = torch.ones(F, device='cuda'), torch.zeros(F, device='cuda')
y = torch.empty(Tv, B, F, device='cuda')
invstd, mm, mv = torch.empty(F, device='cuda'), torch.zeros(F, device='cuda'), torch.ones(F, device='cuda')
rng = torch.tensor([7, 1], dtype=torch.int32, device='cuda')
lens = torch.full((B,), T, dtype=torch.int32, device='cuda')
W = ops.round_tf32(torch.randn(2 * H, 4 * H, device='cuda') / (2 * H) ** 0.5)
gates0 = torch.randn(T, B, 4 * H, device='cuda') * 0.5


Watt = ops.round_tf32(torch.randn(3 * H, 4 * H, device='cuda') / (3 * H) ** 0.5)
Wl = ops.round_tf32(torch.randn(2 * H, H, device='cuda') / (2 * H) ** 0.5)
g1 = torch.ones(1, device='cuda')
mlen = torch.full((B,), Tv, dtype=torch.int32, device='cuda')
values = ops.round_tf32(torch.tanh(torch.randn(Tv, B, H, device='cuda')))
keys = ops.round_tf32(torch.randn(Tv, B, H, device='cuda') * 0.3)
gates_att = torch.randn(T, B, 4 * H, device='cuda') * 0.5


class Drop:
    stream, thr_in, thr_state, thr_out = 8, ops.keep_threshold(0.9), ops.keep_threshold(0.9), ops.keep_threshold(0.9)
    rng = rng


def run():
    ops.bn_apply_train_t(x, sums, B * Tv, gamma, beta, 1e-3, 0.99, y, None, invstd, mm, mv)
    ops.dropout(y, rng, 11, ops.keep_threshold(0.9), round_out=True)
    # cross-modal attention layer forward + backward: the backward ends with the outer products (attn_outer_mma_kernel)
    mb = ops.MechBuffers('scaled_luong', values, keys, mlen, Wl, g=g1)
    att = ops.RnnSeq(T, B, H, lens, gates_att.clone(), Watt[H:], [mb], True)
    mb.dkeys, mb.dvalues = torch.zeros_like(keys), torch.zeros_like(values)
    mb.dWl, mb.dg = torch.zeros_like(Wl), torch.zeros(1, device='cuda')
    att.grad_scale = 1024.0
    att.forward()
    att.backward(torch.randn(T, B, H, device='cuda') * 1e-3, torch.zeros(3 * H, 4 * H, device='cuda'))
    plain = ops.RnnSeq(T, B, H, lens, gates0.clone(), W[H:], drop=Drop)
    plain.grad_scale = 1024.0
    plain.forward()
    plain.backward(torch.randn(T, B, H, device='cuda') * 1e-3, torch.zeros(H, 4 * H, device='cuda'))
    torch.cuda.synchronize()


run()
torch.cuda.profiler.start()
run()
torch.cuda.profiler.stop()
