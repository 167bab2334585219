"""Where a step of the two-product attention-LSTM kernel goes: clock64 stamps of thread 0 / CTA 0 at the
synchronisation points of 16 consecutive steps (library built with `make trace`, -DAP4D_TRACE).

  make trace && AVSR_B200_LIB=avsr_tf1_b200/lib/libavsr_b200_trace.so python tools/ap4d_trace.py [Tm ...]
"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from avsr_tf1_b200 import _lib, ops

B, H, T, Dx = 256, 256, 300, 256
NT, NP = 16, 12
SEG = ['top -> recurrent product done (wait mma1)', 'tmem ld + gate activations + bar.sync', 'c / h / masks + st.async {hs, ho}',
       'HBM stores, gx prefetch, key prefetch', 'wait h all-gather', 'score -> softmax -> context (att_fwd_core)',
       'ctx st.async, masks of t+1, wait ctx all-gather', 'issue attention product + wait mma2', 'tmem ld epilogue + bar.sync',
       'a_t combine + st.async + HBM stores + wait a all-gather', 'issue recurrent product (8 MMAs + commit)']


class Drop:
    def __init__(self):
        self.rng = torch.tensor([1234, 5], dtype=torch.int32, device='cuda')
        self.stream = 8
        self.thr_in = self.thr_state = self.thr_out = ops.keep_threshold(0.9)


def run(Tm):
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
    gates = torch.empty(T, B, 4 * H, device='cuda')
    ops.gemm(x.view(T * B, Dx), W[:Dx], gates.view(T * B, 4 * H))
    for _ in range(2):
        mb = ops.MechBuffers('scaled_luong', values, keys, mlen, Wl, g=g)
        rnn = ops.RnnSeq(T, B, H, lens, gates.clone(), W[Dx:], [mb], True, drop=Drop())
        rnn.forward()
    torch.cuda.synchronize()
    lib = _lib.load()
    fn = lib.avsr_debug_ap4d_trace
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
    buf = np.zeros(NT * NP, np.uint64)
    assert fn(buf.ctypes.data, buf.size) == buf.size
    st = buf.reshape(NT, NP).astype(np.int64)
    # segments within a step: stamp k -> k+1; the last one wraps to stamp 0 of the next step
    seg = np.diff(st, axis=1)[:-1]                      # [NT-1, NP-1]
    wrap = (st[1:, 0] - st[:-1, NP - 1])[:, None]       # issue_rec end -> next loop top
    step = st[1:, 0] - st[:-1, 0]
    clk = float(torch.cuda.clock_rate()) if hasattr(torch.cuda, 'clock_rate') else 1965.0
    mhz = clk if clk < 1e5 else clk / 1e3
    print(f'Tm = {Tm}: step = {step.mean():.0f} clk = {step.mean() / mhz:.2f} us at {mhz:.0f} MHz (median over {NT - 1} steps)')
    med = np.median(seg, axis=0)
    for k, name in enumerate(SEG):
        print(f'  {k:2d} -> {k + 1:2d}  {med[k]:7.0f} clk  {med[k] / mhz:5.2f} us  {name}')
    print(f'  11 ->  0  {np.median(wrap):7.0f} clk  {np.median(wrap) / mhz:5.2f} us  loop back')
    fa = lib.avsr_debug_att_trace
    fa.restype = ctypes.c_int
    fa.argtypes = [ctypes.c_void_p, ctypes.c_int]
    ab = np.zeros(16 * 8, np.uint64)
    assert fa(ab.ctypes.data, ab.size) == ab.size
    a = np.median(np.diff(ab.reshape(16, 8).astype(np.int64), axis=1), axis=0)
    names = ['score sweep (keys prefetched) + 9-shuffle reductions', 'value prefetch issue + barrier 1', 'max: smem, 5 shuffles, barrier 2',
             'exp + sum: 5 shuffles, barrier 3', 'normalise, alignments to smem / HBM, barrier 4', 'context sweep',
             'partials to smem, barrier 5, warp 0 sums']
    print('  inside att_fwd_core:')
    for k, nm in enumerate(names):
        print(f'     a{k} -> a{k + 1}  {a[k]:7.0f} clk  {a[k] / mhz:5.2f} us  {nm}')


if __name__ == '__main__':
    for tm in ([int(a) for a in sys.argv[1:]] or [75, 4, 300]):
        run(tm)
