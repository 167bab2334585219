"""Thin Python wrappers over the C ABI.  torch tensors are used ONLY as device
buffers (allocation, streams); every arithmetic kernel is in libavsr_b200.so."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import AvsrSampling, ATTN_KINDS, AvsrAttnMech, AvsrRnnSeq, check


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _chk_f32(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda or t.dtype != torch.float32:
            raise _lib.AvsrError('expected a CUDA float32 tensor, got %s %s' % (t.device, t.dtype))


KERNEL_CLASSES = ('attn_lstm_fwd', 'attn_lstm_bwd', 'lstm_fwd', 'lstm_bwd', 'gemm')


def kernel_timing(enable: bool) -> bool:
    """Switch the in-library CUDA-event timing of the hot kernels on/off (resets the record)."""
    return bool(_lib.load().avsr_kernel_timing(int(bool(enable))))


def kernel_times():
    """{class: (summed ms, launches)} since kernel_timing(True); synchronises on the recorded events."""
    ms = (C.c_float * 5)()
    n = (C.c_int * 5)()
    check(_lib.load().avsr_kernel_times(ms, n))
    return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(KERNEL_CLASSES)}


def launch_count() -> int:
    return int(_lib.load().avsr_launch_count())


def set_tensor_cores(enable: bool) -> bool:
    return bool(_lib.load().avsr_set_tensor_cores(1 if enable else 0))


def tensor_cores_enabled() -> bool:
    return bool(_lib.load().avsr_get_tensor_cores())


def round_tf32(src: torch.Tensor, dst: torch.Tensor = None) -> torch.Tensor:
    """Round-to-nearest fp32 -> tf32 (the operand format of the tcgen05 products)."""
    src = src.contiguous()
    if dst is None:
        dst = torch.empty_like(src)
    check(_lib.load().avsr_round_tf32(_stream(), src.data_ptr(), dst.data_ptr(), src.numel()))
    return dst


def u8_to_f32(x: torch.Tensor, y: torch.Tensor = None, scale=1.0 / 128.0, shift=-128.0) -> torch.Tensor:
    """Stored lip-crop pixels (uint8) -> the reference's float features (v - 128) / 128 (dataset_writer.py:537)."""
    assert x.dtype == torch.uint8 and x.is_cuda and x.is_contiguous()
    if y is None:
        y = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    check(_lib.load().avsr_u8_to_f32(_stream(), x.data_ptr(), x.numel(), float(scale), float(shift), y.data_ptr()))
    return y


def empty(*shape, dtype=torch.float32):
    return torch.empty(*shape, dtype=dtype, device='cuda')


def zeros(*shape, dtype=torch.float32):
    return torch.zeros(*shape, dtype=dtype, device='cuda')


def gemm(A: torch.Tensor, B: torch.Tensor, out: torch.Tensor, ta=False, tb=False, beta=0.0, bias=None,
         round_out=False):
    """out[M,N] = beta*out + op(A) op(B) (+bias).  2-D tensors, unit inner stride, any row stride."""
    _chk_f32(A, B, out, bias)
    assert A.dim() == 2 and B.dim() == 2 and out.dim() == 2
    assert A.stride(1) == 1 and B.stride(1) == 1 and out.stride(1) == 1, 'inner stride must be 1'
    M, K = (A.shape[1], A.shape[0]) if ta else A.shape
    K2, N = (B.shape[1], B.shape[0]) if tb else B.shape
    assert K == K2, f'gemm inner dims differ: {K} vs {K2}'
    assert tuple(out.shape) == (M, N), f'gemm out shape {tuple(out.shape)} != {(M, N)}'
    check(_lib.load().avsr_gemm(_stream(), int(ta), int(tb), M, N, K, A.data_ptr(), A.stride(0), B.data_ptr(),
                                B.stride(0), out.data_ptr(), out.stride(0), float(beta), _p(bias), int(round_out)))
    return out


def colsum(X: torch.Tensor, out: torch.Tensor):
    """out[N] += column sums of X[M,N]."""
    _chk_f32(X, out)
    check(_lib.load().avsr_colsum(_stream(), X.data_ptr(), X.shape[0], X.shape[1], X.stride(0), out.data_ptr()))


def bn_stats(x2d, sums):
    check(_lib.load().avsr_bn_stats(_stream(), x2d.data_ptr(), x2d.shape[0], x2d.shape[1], sums.data_ptr()))


def bn_apply_train(x2d, sums, count, gamma, beta, eps, momentum, y, xhat, invstd, mm, mv):
    check(_lib.load().avsr_bn_apply_train(_stream(), x2d.data_ptr(), x2d.shape[0], x2d.shape[1], sums.data_ptr(),
                                          float(count), gamma.data_ptr(), beta.data_ptr(), eps, momentum,
                                          y.data_ptr(), _p(xhat), invstd.data_ptr(), _p(mm), _p(mv)))


def bn_apply_train_t(x3d, sums, count, gamma, beta, eps, momentum, y, xhat, invstd, mm, mv):
    """x3d [d0,d1,F] (batch-major) -> y, xhat [d1,d0,F] (frame-major): boundary transpose fused into the BN."""
    d0, d1, F = x3d.shape
    check(_lib.load().avsr_bn_apply_train_t(_stream(), x3d.data_ptr(), d0, d1, F, sums.data_ptr(), float(count),
                                            gamma.data_ptr(), beta.data_ptr(), eps, momentum, y.data_ptr(),
                                            _p(xhat), invstd.data_ptr(), _p(mm), _p(mv)))


def bn_input_grads(Wx, dWx, colsum_dZ, gamma, beta, dgamma, dbeta):
    """dgamma / dbeta of the input BN from the layer-0 weight gradient (see avsr_bn_input_grads); accumulates."""
    F, N = Wx.shape
    check(_lib.load().avsr_bn_input_grads(_stream(), Wx.data_ptr(), Wx.stride(0), dWx.data_ptr(), dWx.stride(0),
                                          colsum_dZ.data_ptr(), gamma.data_ptr(), beta.data_ptr(), F, N,
                                          dgamma.data_ptr(), dbeta.data_ptr()))


def bn_apply_eval(x2d, gamma, beta, mm, mv, eps, y):
    check(_lib.load().avsr_bn_apply_eval(_stream(), x2d.data_ptr(), x2d.shape[0], x2d.shape[1], gamma.data_ptr(),
                                         beta.data_ptr(), mm.data_ptr(), mv.data_ptr(), eps, y.data_ptr()))


def bn_bwd_stats(dy2d, xhat2d, sums2):
    check(_lib.load().avsr_bn_bwd_stats(_stream(), dy2d.data_ptr(), xhat2d.data_ptr(), dy2d.shape[0], dy2d.shape[1],
                                        sums2.data_ptr()))


def bn_bwd_apply(dy2d, xhat2d, sums2, count, gamma, invstd, dx, dgamma, dbeta):
    check(_lib.load().avsr_bn_bwd_apply(_stream(), dy2d.data_ptr(), xhat2d.data_ptr(), dy2d.shape[0], dy2d.shape[1],
                                        sums2.data_ptr(), float(count), gamma.data_ptr(), invstd.data_ptr(),
                                        dx.data_ptr(), _p(dgamma), _p(dbeta)))


def reverse_sequence(x: torch.Tensor, lens: torch.Tensor) -> torch.Tensor:
    """x [T,B,F] -> tf.reverse_sequence along time."""
    T, B, F = x.shape
    y = torch.empty_like(x)
    check(_lib.load().avsr_reverse_sequence(_stream(), x.data_ptr(), y.data_ptr(), T, B, F, lens.data_ptr()))
    return y


def transpose01(x: torch.Tensor) -> torch.Tensor:
    """[d0,d1,F] -> [d1,d0,F] (batch-major <-> frame-major)."""
    x = x.contiguous()
    d0, d1 = x.shape[0], x.shape[1]
    F = int(x.numel() // max(1, d0 * d1))
    y = torch.empty((d1, d0) + tuple(x.shape[2:]), dtype=x.dtype, device=x.device)
    _chk_f32(x)
    check(_lib.load().avsr_transpose01(_stream(), x.data_ptr(), y.data_ptr(), d0, d1, F))
    return y


def embedding_fwd(table, ids, out):
    check(_lib.load().avsr_embedding_fwd(_stream(), table.data_ptr(), table.shape[0], table.shape[1], ids.data_ptr(),
                                         ids.numel(), out.data_ptr()))


def embedding_bwd(dout2d, ids, dtable):
    check(_lib.load().avsr_embedding_bwd(_stream(), dout2d.data_ptr(), ids.data_ptr(), ids.numel(), dtable.shape[0],
                                         dtable.shape[1], dtable.data_ptr()))


def _dev_scalar(x):
    """float -> 1-element device tensor (tests / eager use); tensors pass through."""
    if torch.is_tensor(x):
        return x
    return torch.tensor([float(x)], dtype=torch.float32, device='cuda')


def seq_loss(logits, labels, labels_len, inv_denom, loss_sum, dlogits, label_smoothing=0.0):
    """inv_denom: device scalar tensor (or a float, copied to the device).  label_smoothing > 0: the reference's smoothed,
    UNMASKED mean (include/avsr_b200.h); inv_denom is then 1 / (T*B)."""
    T, B, V = logits.shape
    inv = _dev_scalar(inv_denom)
    check(_lib.load().avsr_seq_loss(_stream(), logits.data_ptr(), T, B, V, labels.data_ptr(), labels.stride(0),
                                    labels_len.data_ptr(), inv.data_ptr(), float(label_smoothing), loss_sum.data_ptr(),
                                    dlogits.data_ptr()))


LOSS_FUNS = {'mc_loss': 1, 'focal_loss': 2}


def seq_loss_devel(logits, labels, labels_len, inv_denom, loss_sum, dlogits, loss_fun, gamma=2.0):
    """devel.py's mc_loss / focal_loss under sequence_loss (include/avsr_b200.h avsr_seq_loss_devel)."""
    T, B, V = logits.shape
    inv = _dev_scalar(inv_denom)
    check(_lib.load().avsr_seq_loss_devel(_stream(), logits.data_ptr(), T, B, V, labels.data_ptr(), labels.stride(0),
                                          labels_len.data_ptr(), inv.data_ptr(), LOSS_FUNS[loss_fun], float(gamma),
                                          loss_sum.data_ptr(), dlogits.data_ptr()))


def au_loss(z, aus, lens, scale_dev, loss_sum, dz):
    """Action-Unit head loss (include/avsr_b200.h avsr_au_loss): z [T,B,2], aus [B,T,2]."""
    T, B, _ = z.shape
    check(_lib.load().avsr_au_loss(_stream(), z.data_ptr(), T, B, aus.data_ptr(), lens.data_ptr(),
                                   scale_dev.data_ptr(), loss_sum.data_ptr(), dz.data_ptr()))


def sumsq(x, out):
    check(_lib.load().avsr_sumsq(_stream(), x.data_ptr(), x.numel(), out.data_ptr()))


def axpy(a, x, y):
    check(_lib.load().avsr_axpy(_stream(), float(a), x.data_ptr(), y.data_ptr(), x.numel()))


def adam_clip_step(params, grads, m, v, sumsq_dev, clip_norm, lr_t, beta1=0.9, beta2=0.999, eps=1e-8,
                   params_tf32=None):
    """lr_t: device scalar tensor (or a float, copied to the device)."""
    lr = _dev_scalar(lr_t)
    check(_lib.load().avsr_adam_clip_step(_stream(), params.data_ptr(), grads.data_ptr(), m.data_ptr(), v.data_ptr(),
                                          params.numel(), sumsq_dev.data_ptr(), float(clip_norm), lr.data_ptr(),
                                          beta1, beta2, eps, _p(params_tf32)))


def same_padding(size, k, stride):
    """TF `SAME`: (output size, pad before, pad after); the extra pixel of an odd total goes to the end."""
    out = -(-size // stride)
    total = max((out - 1) * stride + k - size, 0)
    return out, total // 2, total - total // 2


def conv_geometry(H, W, kh, kw, stride, padding):
    if padding == 'SAME':
        Ho, pt, _ = same_padding(H, kh, stride)
        Wo, pl, _ = same_padding(W, kw, stride)
        return Ho, Wo, pt, pl
    return (H - kh) // stride + 1, (W - kw) // stride + 1, 0, 0


def im2col(x, kh, kw, stride, padding, round_out=True):
    """x [N,H,W,C] -> (cols [N*Ho*Wo, kh*kw*C], geometry)."""
    _chk_f32(x)
    N, H, W, Cc = x.shape
    Ho, Wo, pt, pl = conv_geometry(H, W, kh, kw, stride, padding)
    cols = empty(N * Ho * Wo, kh * kw * Cc)
    check(_lib.load().avsr_im2col(_stream(), x.data_ptr(), N, H, W, Cc, kh, kw, stride, pt, pl, Ho, Wo,
                                  int(bool(round_out)), cols.data_ptr()))
    return cols, (N, H, W, Cc, kh, kw, stride, pt, pl, Ho, Wo)


def col2im(dcols, geom):
    N, H, W, Cc, kh, kw, stride, pt, pl, Ho, Wo = geom
    dx = empty(N, H, W, Cc)
    check(_lib.load().avsr_col2im(_stream(), dcols.data_ptr(), N, H, W, Cc, kh, kw, stride, pt, pl, Ho, Wo,
                                  dx.data_ptr()))
    return dx


def conv2d_direct(x, wmat, bias, kh, kw, stride, padding):
    """Direct NHWC convolution for 8 / 16 output channels (avsr_conv2d_direct).  wmat [kh*kw*Ci, Co]."""
    _chk_f32(x, wmat, bias)
    N, H, W, Ci = x.shape
    Co = wmat.shape[1]
    Ho, Wo, pt, pl = conv_geometry(H, W, kh, kw, stride, padding)
    y = empty(N, Ho, Wo, Co)
    check(_lib.load().avsr_conv2d_direct(_stream(), x.data_ptr(), N, H, W, Ci, wmat.data_ptr(), _p(bias), kh, kw, stride,
                                         pt, pl, Ho, Wo, Co, y.data_ptr()))
    return y


def conv2d_wgrad(x, dy, kh, kw, stride, padding, dW):
    """dW [kh*kw*Ci, Co] += patches(x)^T dy (avsr_conv2d_wgrad)."""
    N, H, W, Ci = x.shape
    Ho, Wo, pt, pl = conv_geometry(H, W, kh, kw, stride, padding)
    Co = dy.shape[-1]
    check(_lib.load().avsr_conv2d_wgrad(_stream(), x.data_ptr(), dy.data_ptr(), N, H, W, Ci, kh, kw, stride, pt, pl, Ho,
                                        Wo, Co, dW.data_ptr()))


def conv2d_tc_supported(Ci, Co, kh, kw, stride):
    return bool(_lib.load().avsr_conv2d_tc_supported(int(Ci), int(Co), int(kh), int(kw), int(stride)))


def conv2d_tc(x, wmat, bias, kh, kw, stride, pad_top, pad_left, Ho, Wo, in_dilation=1, in_bn=None, residual=None, res_bn=None,
              mask_u=None, mask_bn=None, stats=None, out=None):
    """Tensor-core NHWC convolution (avsr_conv2d_tc): x [N,H,W,Ci], wmat [kh*kw*Ci, Co] -> [N,Ho,Wo,Co]; explicit padding /
    output size so that the same call serves as the input gradient (in_dilation = 2: zero-stuffed x).  in_bn / res_bn /
    mask_u + mask_bn / stats: the batch_norm_relu layers around the convolution, fused (see include/avsr_b200.h)."""
    _chk_f32(x, wmat, bias, residual, stats, in_bn, res_bn, mask_u, mask_bn)
    N, H, W, Ci = x.shape
    Co = wmat.shape[1]
    y = empty(N, Ho, Wo, Co) if out is None else out
    check(_lib.load().avsr_conv2d_tc(_stream(), x.data_ptr(), N, H, W, Ci, wmat.data_ptr(), _p(bias), kh, kw, stride, pad_top,
                                     pad_left, Ho, Wo, Co, in_dilation, _p(in_bn), _p(residual), _p(res_bn), _p(mask_u),
                                     _p(mask_bn), _p(stats), y.data_ptr()))
    return y


def conv2d_wgrad_tc(x, dy, kh, kw, stride, padding, dW, in_bn=None, dbias=None):
    """dW [kh*kw*Ci, Co] += patches(x')^T dy on tensor cores (avsr_conv2d_wgrad_tc); x' = relu(x scale + shift) with in_bn."""
    _chk_f32(x, dy, dW, in_bn, dbias)
    N, H, W, Ci = x.shape
    Ho, Wo, pt, pl = conv_geometry(H, W, kh, kw, stride, padding)
    Co = dy.shape[-1]
    check(_lib.load().avsr_conv2d_wgrad_tc(_stream(), x.data_ptr(), _p(in_bn), dy.data_ptr(), N, H, W, Ci, kh, kw, stride, pt,
                                           pl, Ho, Wo, Co, dW.data_ptr(), _p(dbias)))


def bn_finalize(sums, count, gamma, beta, eps, momentum, moving_mean, moving_var):
    """coef [4C] = (scale, shift, invstd, -mean invstd) of a batch_norm_relu from fused (sum, sum of squares) statistics."""
    C = gamma.numel()
    coef = empty(4 * C)
    check(_lib.load().avsr_bn_finalize(_stream(), sums.data_ptr(), float(count), gamma.data_ptr(), beta.data_ptr(), eps,
                                       momentum, C, _p(moving_mean), _p(moving_var), coef.data_ptr()))
    return coef


def bn_coef_eval(gamma, beta, moving_mean, moving_var, eps):
    C = gamma.numel()
    coef = empty(4 * C)
    check(_lib.load().avsr_bn_coef_eval(_stream(), gamma.data_ptr(), beta.data_ptr(), moving_mean.data_ptr(),
                                        moving_var.data_ptr(), eps, C, coef.data_ptr()))
    return coef


def bn_relu_bwd_apply(d, u, coef, sums2, count, residual=None, out=None):
    """du of a batch_norm_relu from the ReLU-masked gradient d, the saved BN input u and sums2 = (sum d, sum d xhat)."""
    _chk_f32(d, u, coef, sums2, residual)
    C = d.shape[-1]
    du = torch.empty_like(d) if out is None else out
    check(_lib.load().avsr_bn_relu_bwd_apply(_stream(), d.data_ptr(), u.data_ptr(), coef.data_ptr(), sums2.data_ptr(),
                                             float(count), _p(residual), d.numel() // C, C, du.data_ptr()))
    return du


def selu_fwd(x, out=None):
    y = torch.empty_like(x) if out is None else out
    check(_lib.load().avsr_selu_fwd(_stream(), x.data_ptr(), x.numel(), y.data_ptr()))
    return y


def selu_bwd(y, dy, out=None):
    dx = torch.empty_like(dy) if out is None else out
    check(_lib.load().avsr_selu_bwd(_stream(), y.data_ptr(), dy.data_ptr(), dy.numel(), dx.data_ptr()))
    return dx


def highway_fwd(x, pre, out, want_operand=True):
    """HighwayWrapper (cells.py:89-90): y = x * sigmoid(pre) + out * (1 - sigmoid(pre)); returns (y, operand copy of y)."""
    y = torch.empty_like(x)
    y_op = torch.empty_like(x) if want_operand else None
    check(_lib.load().avsr_highway_fwd(_stream(), x.data_ptr(), pre.data_ptr(), out.data_ptr(), x.numel(), y.data_ptr(),
                                       _p(y_op)))
    return y, y_op


def highway_bwd(dy, x, pre, out):
    """Returns (dx through the carried input, dout, dpre)."""
    dx, dout, dpre = torch.empty_like(dy), torch.empty_like(dy), torch.empty_like(dy)
    check(_lib.load().avsr_highway_bwd(_stream(), dy.data_ptr(), x.data_ptr(), pre.data_ptr(), out.data_ptr(), dy.numel(),
                                       dx.data_ptr(), dout.data_ptr(), dpre.data_ptr()))
    return dx, dout, dpre


def relu_fwd(x, out=None):
    y = torch.empty_like(x) if out is None else out
    check(_lib.load().avsr_relu_fwd(_stream(), x.data_ptr(), x.numel(), y.data_ptr()))
    return y


def relu_bwd(y, dy, out=None):
    dx = torch.empty_like(dy) if out is None else out
    check(_lib.load().avsr_relu_bwd(_stream(), y.data_ptr(), dy.data_ptr(), dy.numel(), dx.data_ptr()))
    return dx


OPTIMISERS = {'Adam': 0, 'Nadam': 1, 'AdamW': 2, 'Momentum': 3}


def optim_clip_step(kind, params, grads, m, v, sumsq_dev, clip_norm, lr_t, beta1=0.9, beta2=0.999, eps=1e-8,
                    weight_decay=0.0, params_tf32=None):
    """clip_by_global_norm + one of the reference's optimisers (include/avsr_b200.h avsr_optim_clip_step)."""
    lr = _dev_scalar(lr_t)
    check(_lib.load().avsr_optim_clip_step(_stream(), OPTIMISERS[kind], params.data_ptr(), grads.data_ptr(),
                                           m.data_ptr(), v.data_ptr(), params.numel(), sumsq_dev.data_ptr(),
                                           float(clip_norm), lr.data_ptr(), beta1, beta2, eps, float(weight_decay),
                                           _p(params_tf32)))


def normed_v_fwd(v, g, veff):
    check(_lib.load().avsr_normed_v_fwd(_stream(), v.data_ptr(), g.data_ptr(), v.numel(), veff.data_ptr()))


def normed_v_bwd(v, g, dveff, dv, dg):
    check(_lib.load().avsr_normed_v_bwd(_stream(), v.data_ptr(), g.data_ptr(), dveff.data_ptr(), v.numel(),
                                        dv.data_ptr(), dg.data_ptr()))


def keep_threshold(keep_prob) -> int:
    """keep probability -> the uint32 threshold of the counter-based generator (0 = never drop)."""
    keep_prob = float(keep_prob)
    if keep_prob >= 1.0:
        return 0
    if not keep_prob > 0.0:
        raise ValueError('keep probability must be in (0, 1]')
    return max(1, min(int(keep_prob * 4294967296.0), 4294967295))


def dropout(x, rng, stream_id, thr, round_out=False, out=None, first=0):
    """tf.nn.dropout with the library's generator (include/avsr_b200.h avsr_dropout); the same call maps dy -> dx.
    `first`: mask index of x's first element (a slice of a larger masked tensor)."""
    _chk_f32(x)
    y = torch.empty_like(x) if out is None else out
    check(_lib.load().avsr_dropout(_stream(), x.data_ptr(), x.numel(), int(first), rng.data_ptr(), int(stream_id),
                                   int(thr), int(bool(round_out)), y.data_ptr()))
    return y


def sched_sample(logits, rng, stream_id, t, thr_p, true_next, next_ids, sampled):
    B, V = logits.shape
    check(_lib.load().avsr_sched_sample(_stream(), logits.data_ptr(), B, V, rng.data_ptr(), int(stream_id), int(t),
                                        int(thr_p), true_next.data_ptr(), next_ids.data_ptr(), sampled.data_ptr()))


def greedy_pick(logits, eos, finished, sample_out, next_ids):
    B, V = logits.shape
    check(_lib.load().avsr_greedy_pick(_stream(), logits.data_ptr(), B, V, eos, finished.data_ptr(),
                                       sample_out.data_ptr(), next_ids.data_ptr()))


def beam_step(logits, B, W, eos, lpw, log_probs, finished, lengths, word, parent, score):
    V = logits.shape[1]
    check(_lib.load().avsr_beam_step(_stream(), logits.data_ptr(), B, W, V, eos, float(lpw), log_probs.data_ptr(),
                                     finished.data_ptr(), lengths.data_ptr(), word.data_ptr(), parent.data_ptr(),
                                     score.data_ptr()))


def gather_rows(src2d, idx, dst2d):
    check(_lib.load().avsr_gather_rows(_stream(), src2d.data_ptr(), idx.data_ptr(), idx.numel(), src2d.shape[1],
                                       dst2d.data_ptr()))


# --------------------------------------------------------------------------- #
# recurrent sequence op
# --------------------------------------------------------------------------- #
class MechBuffers:
    """Device buffers of one attention mechanism for one sequence call."""

    def __init__(self, kind: str, values, keys, mem_len, Wl, Wq=None, v=None, g=None, bias=None):
        self.kind = kind
        self.values, self.keys, self.mem_len = values, keys, mem_len
        self.Wl, self.Wq, self.v, self.g, self.bias = Wl, Wq, v, g, bias
        self.Tm, self.B, self.Dm = values.shape
        self.A = keys.shape[2]
        self.align = self.hc = self.pq = None
        self.dkeys = self.dvalues = self.dWl = self.dWq = self.dv = self.dg = self.dbias = self.dpq = None
        self.ds = self.dhc = None
        self.values_op = None  # tf32-rounded copy of the memory (operand of tensor-core products), or None = values

    def fill(self, m: AvsrAttnMech):
        if self.values_op is not None and not self.values_op.is_contiguous():
            # e.g. the h columns of a layer's state rows: the library reads a dense [Tm*B, Dm] operand
            self.values_op = self.values_op.contiguous()
        m.kind = ATTN_KINDS[self.kind]
        m.Tm, m.Dm, m.A = self.Tm, self.Dm, self.A
        for k in ('values', 'keys', 'mem_len', 'Wl', 'Wq', 'v', 'g', 'bias', 'align', 'hc', 'pq', 'dkeys', 'dvalues',
                  'dWl', 'dWq', 'dv', 'dg', 'dbias', 'dpq', 'ds', 'dhc', 'values_op'):
            setattr(m, k, _p(getattr(self, k)))


class Sampling:
    """ScheduledEmbeddingTrainingHelper inside the recurrence (AvsrSampling in include/avsr_b200.h): device tensors the
    recurrent op reads / updates when it draws the decoder inputs itself."""

    def __init__(self, Wd, bd, embedding, Wx, bias, used_ids, sample_ids, x, stream, thr_p):
        self.Wd, self.bd, self.embedding, self.Wx, self.bias = Wd, bd, embedding, Wx, bias
        self.used_ids, self.sample_ids, self.x = used_ids, sample_ids, x
        self.V, self.E = int(embedding.shape[0]), int(embedding.shape[1])
        self.stream, self.thr_p = int(stream), int(thr_p)
        for t in (Wd, embedding, Wx, x):
            _chk_f32(t)

    def struct(self) -> AvsrSampling:
        s = AvsrSampling()
        for k in ('Wd', 'bd', 'embedding', 'Wx', 'bias', 'used_ids', 'sample_ids', 'x'):
            setattr(s, k, _p(getattr(self, k)))
        s.V, s.E, s.stream, s.thr_p = self.V, self.E, self.stream, self.thr_p
        return s


class RnnSeq:
    """One dynamic_rnn / dynamic_decode loop (see AvsrRnnSeq in include/avsr_b200.h)."""

    def __init__(self, T, B, H, lens, gates, Wrec, mechs: Sequence[MechBuffers] = (), output_attention=False,
                 c0=None, h0=None, s0=None, drop=None):
        self.T, self.B, self.H = T, B, H
        self.drop = drop  # DropState (layers.py) or None: DropoutWrapper masks applied inside the loop
        self.stepwise = False
        self.sampling = None  # Sampling or None: scheduled sampling inside a whole-sequence forward
        self.rng = None       # generator words when there is no DropState (sampling without dropout)
        self.lens, self.gates, self.Wrec, self.c0 = lens, gates, Wrec, c0
        self.mechs = list(mechs)
        self.oa = bool(output_attention) and len(self.mechs) > 0
        self.At = sum(m.A for m in self.mechs)
        SW = self.At + H
        self.S = empty(T + 1, B, SW)
        if s0 is not None:  # full state row [attention | h] (step-wise decoding)
            self.S[0].copy_(s0)
        else:
            if self.At > 0:
                self.S[0, :, :self.At].zero_()
            if h0 is None:
                self.S[0, :, self.At:].zero_()
            else:
                self.S[0, :, self.At:].copy_(h0)
        self.craw = empty(T, B, H)
        self.out = empty(T, B, self.At if self.oa else H)
        self.cT = empty(B, H)
        self.hT = empty(B, H)
        for m in self.mechs:
            m.align = empty(T, B, m.Tm)
            m.hc = empty(T, B, H + m.Dm)
            if 'bahdanau' in m.kind:
                m.pq = empty(T, B, m.A)
        maxHD = max([H + m.Dm for m in self.mechs], default=0)
        maxA = max([m.A for m in self.mechs], default=0)
        maxTm = max([m.Tm for m in self.mechs], default=0)
        nwork = int(_lib.load().avsr_rnn_work_floats(B, H, self.At, maxHD, maxA, maxTm))
        self.work = empty(max(nwork, 4))
        self.dZ = self.dA = None
        # power-of-two scale that brings the gate gradients near 1 before they enter the tensor core as fp16
        # (persistent attention backward); losses averaged over n tokens have gradients ~ 1/n
        self.grad_scale = 1.0

    def _desc(self, **bw) -> AvsrRnnSeq:
        r = AvsrRnnSeq()
        r.T, r.B, r.H, r.n_mech, r.output_attention = self.T, self.B, self.H, len(self.mechs), int(self.oa)
        r.len, r.gates, r.Wrec, r.c0 = _p(self.lens), _p(self.gates), _p(self.Wrec), _p(self.c0)
        r.S, r.craw, r.out, r.cT, r.hT = _p(self.S), _p(self.craw), _p(self.out), _p(self.cT), _p(self.hT)
        for k, m in enumerate(self.mechs):
            m.fill(r.mech[k])
        r.work = _p(self.work)
        r.grad_scale = float(self.grad_scale)
        if self.drop is not None:
            d = self.drop
            r.rng, r.drop_stream = _p(d.rng), d.stream
            r.thr_in, r.thr_state, r.thr_out = d.thr_in, d.thr_state, d.thr_out
        elif self.rng is not None:
            r.rng = _p(self.rng)
        r.stepwise = int(self.stepwise)
        if self.sampling is not None:
            self._samp_struct = self.sampling.struct()  # kept alive for the duration of the call
            r.samp = C.pointer(self._samp_struct)
        for k, v in bw.items():
            setattr(r, k, _p(v))
        return r

    def forward(self):
        r = self._desc()
        check(_lib.load().avsr_rnn_seq_fwd(_stream(), C.byref(r)))
        return self.out

    def sampling_fused(self) -> bool:
        """True if a whole-sequence forward of this layer can draw scheduled samples inside the recurrence."""
        return bool(_lib.load().avsr_rnn_sampling_fused(C.byref(self._desc())))

    def forward_range(self, t_begin, t_end):
        """Steps [t_begin, t_end) only (step-wise kernels): scheduled sampling interleaves its draws with the
        recurrence.  The backward pass of such a forward runs step-wise too."""
        self.stepwise = True
        r = self._desc()
        r.t_begin, r.t_end = int(t_begin), int(t_end)
        check(_lib.load().avsr_rnn_seq_fwd(_stream(), C.byref(r)))
        return self.out

    def backward(self, dout, dWrec, dcT=None, dhT=None, want_init_grad=False, dbias=None):
        """Fills self.dZ [T,B,4H]; accumulates dWrec, the mechanism grads (set on the MechBuffers) and, if given,
        the bias gradient dbias [4H] += column sums of dZ."""
        T, B, H = self.T, self.B, self.H
        self.dZ = empty(T, B, 4 * H)
        if self.At > 0:
            self.dA = empty(T, B, self.At)
        self.dc0 = empty(B, H) if want_init_grad else None
        self.dh0 = empty(B, H) if want_init_grad else None
        for m in self.mechs:
            if 'bahdanau' in m.kind:
                m.dpq = empty(T, B, m.A)
            m.ds = empty(T, B, m.Tm)
            m.dhc = empty(T, B, H + m.Dm)
        r = self._desc(dout=dout, dcT=dcT, dhT=dhT, dZ=self.dZ, dA=self.dA, dWrec=dWrec, dc0=self.dc0, dh0=self.dh0,
                       dbias=dbias)
        check(_lib.load().avsr_rnn_seq_bwd(_stream(), C.byref(r)))
        return self.dZ
