"""Diagnostic: error growth over time of the attention-LSTM layer under dropout, persistent vs step-wise (tf32 mode)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import avsr_oracle as O
from tests import test_gpu_ops as G

ops = G.ops_mod()
kinds, B, T, Dx, H, Tms, Dms, keep = ('scaled_luong',), 8, 24, 80, 256, (96,), (256,), (0.9, 0.85, 0.95)
x, lens, W, b, specs, c0, h0, rng = G._attn_case(kinds, B, T, Dx, H, Tms, Dms, sum(Tms) + B)
f64 = lambda a: a.astype(np.float64)
words, stream = (4321, 17), 12
odrop = O.DropSpec(words, stream, keep)
r = O.attn_rnn_fwd(f64(x), lens, f64(W), f64(b), [G._spec64(s) for s in specs], init_cell=(f64(c0), f64(h0)), drop=odrop)
res = {}
for mode in ('persist', 'stepwise'):
    if mode == 'stepwise':
        os.environ['AVSR_NO_ATTN_PERSIST'] = '1'
    ops.set_tensor_cores(True)
    drop = G._Drop(ops, words, stream, keep)
    dev, opnd = G.dev, G.opnd
    xt = ops.dropout(dev(x.transpose(1, 0, 2)).contiguous(), drop.rng, drop.stream + 3, drop.thr_in, round_out=True)
    Wd, bd, ld = opnd(dev(W), True), dev(b), dev(lens, torch.int32)
    gates = torch.empty(T, B, 4 * H, device='cuda')
    ops.gemm(xt.view(T * B, Dx), Wd[:Dx], gates.view(T * B, 4 * H), bias=bd)
    s = specs[0]
    Tm, Dm, A = s.memory.shape[1], s.memory.shape[2], H
    values = dev(s.memory.transpose(1, 0, 2))
    keys = torch.empty(Tm, B, A, device='cuda')
    ops.gemm(opnd(values, True).view(Tm * B, Dm), opnd(dev(s.Wm), True), keys.view(Tm * B, A))
    g = dev(np.asarray(s.g).reshape(1))
    mb = ops.MechBuffers(s.kind, values, keys, dev(s.mem_len, torch.int32), opnd(dev(s.Wl), True), g=g)
    rnn = ops.RnnSeq(T, B, H, ld, gates, Wd[Dx:], [mb], True, c0=dev(c0), h0=dev(h0), drop=drop)
    out = rnn.forward().transpose(0, 1).cpu().numpy().astype(np.float64)
    res[mode] = out
    scale = np.abs(r['outputs']).max()
    err_t = np.abs(out - r['outputs']).max(axis=(0, 2)) / scale
    print(mode, 'max scaled error per step:', np.array2string(err_t, precision=4, max_line_width=200))
d = np.abs(res['persist'] - res['stepwise']).max(axis=(0, 2)) / np.abs(r['outputs']).max()
print('persist vs stepwise per step:', np.array2string(d, precision=4, max_line_width=200))
print('lens', lens)
