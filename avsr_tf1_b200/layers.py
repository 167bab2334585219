"""Compute blocks of the hot path, each with an explicit forward and backward
(there is no autograd: every kernel is hand-written, see csrc/).  All sequences
are frame-major [T, B, F] device tensors."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import ops
from .params import ParamSpec

BN_EPS = 1e-3  # tf.layers.batch_normalization defaults (encoder.py:45)
BN_MOMENTUM = 0.99


class BuildContext:
    """Collects ParamSpecs while the model is being assembled, then owns the store."""

    def __init__(self):
        self.specs: List[ParamSpec] = []
        self.store = None
        self.world_size = 1
        # global batch / this rank's batch: batch-norm statistics are all-reduced SUMS over the global batch, so their
        # row count is rows_local * batch_scale.  Ranks may hold unequal shares of a global batch (a batch that does not
        # divide evenly); Seq2SeqModel sets the exact ratio per batch when the iterator reports the global batch size,
        # else it is world_size (equal shares).  All ranks must pad a stream to the same length (RecordBatcher does).
        self.batch_scale = 1.0
        self.grad_scale = 1.0  # power of two ~ number of target tokens (set per step by Seq2SeqModel)
        self.allreduce = None  # callable(tensor) -> None (in-place sum), set by Seq2SeqModel under DP
        self.rng = None  # device int32[2] {seed, step}: counter-based generator of dropout / scheduled sampling
        # Independent recurrent chains (forward / backward stacks of a BiLSTM encoder, video / audio encoders) run side
        # by side on two streams when two persistent kernels fit the GPU together: a cluster-of-4 kernel occupies
        # 4 * ceil(B / 8) SMs, so up to 144 utterances two of them share the 148 SMs (Seq2SeqModel sets this per batch).
        self.parallel_chains = False
        self._side = {}
        self._streams = 0
        self.streams = {}  # cell name (variable prefix) -> first stream id: lets the oracle regenerate every mask

    def fork(self, level=0):
        """Side stream forked from the current one (also under graph capture); `level` keeps nested forks apart."""
        if level not in self._side:
            self._side[level] = torch.cuda.Stream()
        self._side[level].wait_stream(torch.cuda.current_stream())
        return self._side[level]

    def join(self, level=0):
        torch.cuda.current_stream().wait_stream(self._side[level])

    def new_stream(self, n=4):
        """Reserves n consecutive stream ids of the generator (one independent random sequence each)."""
        base = 8 + self._streams
        self._streams += n
        return base

    def drop_state(self, cell, name):
        """DropState of a cell spec built by cells.build_rnn_layers, or None when its DropoutWrapper is off."""
        if not getattr(cell, 'use_dropout', False):
            return None
        ds = DropState(self, cell.dropout_probability)
        self.streams[name] = ds.stream
        return ds

    def declare(self, name, shape, init, trainable=True):
        if any(s.name == name for s in self.specs):
            raise Exception('duplicate variable ' + name)
        self.specs.append(ParamSpec(name, tuple(int(x) for x in shape), init, trainable))
        return name

    def p(self, name):
        return self.store.p(name)

    def g(self, name):
        return self.store.g(name)

    def w(self, name):
        """Parameter as the operand of a matrix product (tf32-rounded copy in tensor-core mode)."""
        return self.store.w(name, ops.tensor_cores_enabled())


class DropState:
    """DropoutWrapper(input_keep, state_keep, output_keep) of one cell (cells.py:46-54), non-variational.
    Streams: +0 attention part of the cell input, +1 recurrent h, +2 cell output (these three are applied inside the
    recurrent op), +3 the x part of the cell input (applied to the whole sequence before the x-projection)."""

    def __init__(self, ctx: 'BuildContext', keep_probs):
        self.ctx = ctx
        self.stream = ctx.new_stream(4)
        self.thr_in, self.thr_state, self.thr_out = (ops.keep_threshold(p) for p in keep_probs)

    @property
    def any(self):
        return bool(self.thr_in or self.thr_state or self.thr_out)

    @property
    def rng(self):
        return self.ctx.rng

    def drop_input(self, x):
        """x part of the cell input -> product operand (tf32-rounded in tensor-core mode)."""
        if self.thr_in:
            return ops.dropout(x, self.rng, self.stream + 3, self.thr_in, round_out=True)
        return ops.round_tf32(x) if ops.tensor_cores_enabled() else x

    def drop_input_grad(self, dx):
        if self.thr_in and dx is not None:
            ops.dropout(dx, self.rng, self.stream + 3, self.thr_in, out=dx)
        return dx


class BatchNormInput:
    """tf.layers.batch_normalization(axis=-1) on the raw features (encoder.py:44-50):
    batch statistics over every (b, t) position INCLUDING padding; momentum 0.99, eps 1e-3."""

    def __init__(self, ctx: BuildContext, scope: str, F: int):
        self.ctx, self.F = ctx, F
        pre = f'{scope}/batch_normalization/'
        self.gamma = ctx.declare(pre + 'gamma', (F,), 'ones')
        self.beta = ctx.declare(pre + 'beta', (F,), 'zeros')
        self.mm = ctx.declare(pre + 'moving_mean', (F,), 'zeros', trainable=False)
        self.mv = ctx.declare(pre + 'moving_variance', (F,), 'ones', trainable=False)

    def forward(self, x: torch.Tensor, train: bool, batch_major: bool = False, keep_xhat: bool = True) -> torch.Tensor:
        """x [T,B,F] frame-major, or [B,T,F] with batch_major=True (the reference's layout); the result is frame-major
        either way - in train mode the change of layout is fused into the normalisation.  keep_xhat=False: the
        normalised features are not stored; dgamma / dbeta then come from `backward_from_layer0`."""
        ctx = self.ctx
        if batch_major and not train:
            x, batch_major = ops.transpose01(x), False
        d0, d1, F = x.shape
        T, B = (d1, d0) if batch_major else (d0, d1)
        y = ops.empty(T, B, F)
        if not train:
            ops.bn_apply_eval(x.view(T * B, F), ctx.p(self.gamma), ctx.p(self.beta), ctx.p(self.mm), ctx.p(self.mv),
                              BN_EPS, y)
            return y
        sums = ops.zeros(2 * F)
        ops.bn_stats(x.view(T * B, F), sums)  # column sums: the order of the rows does not matter
        count = float(T * B)
        if ctx.world_size > 1:  # exact large-batch statistics under data parallelism
            ctx.allreduce(sums)
            count *= ctx.batch_scale
        self.xhat = ops.empty(T, B, F) if keep_xhat else None
        self.invstd = ops.empty(F)
        self.count = count
        if batch_major:
            ops.bn_apply_train_t(x, sums, count, ctx.p(self.gamma), ctx.p(self.beta), BN_EPS, BN_MOMENTUM, y,
                                 self.xhat, self.invstd, ctx.p(self.mm), ctx.p(self.mv))
        else:
            ops.bn_apply_train(x.view(T * B, F), sums, count, ctx.p(self.gamma), ctx.p(self.beta), BN_EPS, BN_MOMENTUM,
                               y.view(T * B, F), None if self.xhat is None else self.xhat.view(T * B, F), self.invstd,
                               ctx.p(self.mm), ctx.p(self.mv))
        return y

    def backward_from_layer0(self, layer0_ops) -> None:
        """dgamma / dbeta from what the weight-gradient pass of the layer(s) fed by this normalisation has just formed
        (their dWx = y^T dZ and bias gradient colsum(dZ) sit in the gradient buffer, still free of the L2 term and of
        other ranks' sums): no gradient wrt y, no stored xhat.  See avsr_bn_input_grads."""
        ctx = self.ctx
        for op in layer0_ops:
            I = op.I if hasattr(op, 'I') else op.Dx  # rows of the cell kernel that multiply the layer input
            ops.bn_input_grads(ctx.w(op.kernel)[:I], ctx.g(op.kernel)[:I], ctx.g(op.bias), ctx.p(self.gamma),
                               ctx.p(self.beta), ctx.g(self.gamma), ctx.g(self.beta))

    def backward(self, dy: torch.Tensor, need_dx: bool = True) -> Optional[torch.Tensor]:
        """Accumulates dgamma / dbeta; the gradient wrt the raw features is only formed on request (nothing upstream
        of the input normalisation is trained on this path)."""
        ctx = self.ctx
        T, B, F = dy.shape
        if self.xhat is None:
            raise Exception('input BN: forward ran with keep_xhat=False, use backward_from_layer0')
        dy2, xh2 = dy.view(T * B, F), self.xhat.view(T * B, F)
        sums2 = ops.zeros(2 * F)
        ops.bn_bwd_stats(dy2, xh2, sums2)
        local = sums2
        if ctx.world_size > 1 and need_dx:
            local = sums2.clone()  # dgamma/dbeta stay local sums (the gradient all-reduce adds them up)
            ctx.allreduce(sums2)
        dx = None
        if need_dx:
            dx = torch.empty_like(dy)
            ops.bn_bwd_apply(dy2, xh2, sums2, self.count, ctx.p(self.gamma), self.invstd, dx, None, None)
        ops.axpy(1.0, local[F:], ctx.g(self.gamma))
        ops.axpy(1.0, local[:F], ctx.g(self.beta))
        return dx


class InstanceNormInput:
    """tf.contrib.layers.instance_norm(inputs) on the [B,T,F] features (encoder.py:51-55; TF 1.13 defaults: center, scale,
    epsilon 1e-6): moments over the time axis of every (utterance, feature), padded frames included; gamma / beta per
    feature.  Frame-major [T,B,F] is a [T, B*F] matrix whose COLUMN statistics are exactly these moments, so the batch-norm
    kernels serve it with gamma / beta tiled over the utterances (no moving statistics: the same in training and inference;
    nothing to exchange under data parallelism)."""
    EPS = 1e-6

    def __init__(self, ctx: BuildContext, scope: str, F: int):
        self.ctx, self.F = ctx, F
        self.beta = ctx.declare(f'{scope}/InstanceNorm/beta', (F,), 'zeros')
        self.gamma = ctx.declare(f'{scope}/InstanceNorm/gamma', (F,), 'ones')

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x [T,B,F] frame-major -> normalised features, a product operand (tf32-rounded in tensor-core mode)."""
        ctx = self.ctx
        T, B, F = x.shape
        x2 = x.contiguous().view(T, B * F)
        sums = ops.zeros(2 * B * F)
        ops.bn_stats(x2, sums)
        self.gamma_t = ctx.p(self.gamma).repeat(B)
        beta_t = ctx.p(self.beta).repeat(B)
        y = ops.empty(T, B, F)
        self.xhat, self.invstd = ops.empty(T, B * F), ops.empty(B * F)
        ops.bn_apply_train(x2, sums, float(T), self.gamma_t, beta_t, self.EPS, 0.0, y.view(T, B * F), self.xhat, self.invstd,
                           None, None)
        return y

    def backward(self, dy: torch.Tensor) -> torch.Tensor:
        ctx = self.ctx
        T, B, F = dy.shape
        dy2 = dy.contiguous().view(T, B * F)
        sums2 = ops.zeros(2 * B * F)
        ops.bn_bwd_stats(dy2, self.xhat, sums2)
        ops.axpy(1.0, sums2[B * F:].view(B, F).sum(0), ctx.g(self.gamma))
        ops.axpy(1.0, sums2[:B * F].view(B, F).sum(0), ctx.g(self.beta))
        dx = ops.empty(T, B, F)
        ops.bn_bwd_apply(dy2, self.xhat, sums2, float(T), self.gamma_t, self.invstd, dx.view(T, B * F), None, None)
        self.xhat = None
        return dx


class LSTMLayerOp:
    """One LSTMCell layer under dynamic_rnn (cells.py:14-18, encoder.py:80)."""

    def __init__(self, ctx: BuildContext, prefix: str, in_dim: int, H: int, drop: 'DropState' = None,
                 share: 'LSTMLayerOp' = None):
        self.ctx, self.I, self.H = ctx, in_dim, H
        self.drop = drop
        if share is not None:  # encoder_weight_sharing (cells.py:77-78): this position runs with another layer's variables
            if (share.I, share.H) != (in_dim, H):
                raise ValueError('weight sharing needs layers of equal shape')
            self.kernel, self.bias = share.kernel, share.bias
            return
        self.kernel = ctx.declare(prefix + '/kernel', (in_dim + H, 4 * H), 'lstm_kernel')
        self.bias = ctx.declare(prefix + '/bias', (4 * H,), 'zeros')

    def forward(self, x: torch.Tensor, lens: torch.Tensor):
        """x [T,B,I] must be a product operand (tf32-rounded in tensor-core mode).  Returns the exact
        outputs (zero past the length); `self.operand` is the same sequence as the NEXT product's operand
        (rounded h; rows past the length carry the last state instead of zero, which no consumer reads).
        With a DropoutWrapper: x is the un-dropped layer input (any precision), the outputs carry the output
        dropout and `operand` is the outputs (the consumer drops / rounds them itself)."""
        T, B, I = x.shape
        H = self.H
        W, b = self.ctx.w(self.kernel), self.ctx.p(self.bias)
        if self.drop is not None:
            x = self.drop.drop_input(x)
        gates = ops.empty(T, B, 4 * H)
        ops.gemm(x.reshape(T * B, I), W[:I], gates.view(T * B, 4 * H), bias=b)
        self.x = x
        self.rnn = ops.RnnSeq(T, B, H, lens, gates, W[I:], drop=self.drop)
        out = self.rnn.forward()
        self.final = (self.rnn.cT, self.rnn.hT)  # (with state dropout hT is the dropped h, as in the reference)
        self.operand = self.rnn.S[1:] if self.drop is None else out
        return out

    def backward(self, dout, dstate=None, need_dx=True):
        T, B, I = self.x.shape
        H = self.H
        W = self.ctx.w(self.kernel)
        gW, gb = self.ctx.g(self.kernel), self.ctx.g(self.bias)
        dcT, dhT = dstate if dstate is not None else (None, None)
        self.rnn.grad_scale = self.ctx.grad_scale  # fp16 dz operand of the cluster-of-4 backward kernel
        dZ = self.rnn.backward(dout, gW[I:], dcT=dcT, dhT=dhT, dbias=gb)  # (bias gradient inside the recurrent op)
        dZ2 = dZ.view(T * B, 4 * H)
        ops.gemm(self.x.reshape(T * B, I), dZ2, gW[:I], ta=True, beta=1.0)
        dx = None
        if need_dx:
            dx = ops.empty(T, B, I)
            ops.gemm(dZ2, W[:I], dx.view(T * B, I), tb=True)
            if self.drop is not None:
                self.drop.drop_input_grad(dx)
        self.rnn = None
        return dx


class MechDef:
    """Static description of one attention mechanism (attention.py:5-88)."""

    def __init__(self, ctx: BuildContext, kind: str, wrap_prefix: str, idx: int, mem_layer_name: str, H: int, Dm: int,
                 A: int):
        self.kind, self.H, self.Dm, self.A = kind, H, Dm, A
        sfx = '' if idx == 0 else f'_{idx}'
        fam = 'luong_attention' if 'luong' in kind else 'bahdanau_attention'
        self.Wm = ctx.declare(mem_layer_name, (Dm, A), 'glorot')
        self.Wq = self.v = self.g = self.b = None
        if kind == 'scaled_luong':
            self.g = ctx.declare(f'{wrap_prefix}/{fam}{sfx}/attention_g', (), 'ones')
        if 'bahdanau' in kind:
            self.Wq = ctx.declare(f'{wrap_prefix}/{fam}{sfx}/query_layer/kernel', (H, A), 'glorot')
            self.v = ctx.declare(f'{wrap_prefix}/{fam}{sfx}/attention_v', (A,), 'glorot')
        if kind == 'normed_bahdanau':
            self.g = ctx.declare(f'{wrap_prefix}/{fam}{sfx}/attention_g', (), 'const:%r' % (1.0 / A) ** 0.5)
            self.b = ctx.declare(f'{wrap_prefix}/{fam}{sfx}/attention_b', (A,), 'zeros')
        self.Wl = ctx.declare(f'{wrap_prefix}/attention_layer{sfx}/kernel', (H + Dm, A), 'glorot')

    @property
    def output_attention(self):
        return 'luong' in self.kind


class AttnLSTMOp:
    """AttentionWrapper(LSTMCell) under dynamic_rnn / dynamic_decode (attention.py:132-191)."""

    def __init__(self, ctx: BuildContext, wrap_prefix: str, in_dim: int, H: int, mechs: Sequence[MechDef],
                 drop: 'DropState' = None):
        self.ctx, self.Dx, self.H, self.mechs = ctx, in_dim, H, list(mechs)
        self.drop = drop  # DropoutWrapper of the wrapped cell: acts on concat(x, attention), the state h, the cell output
        self.At = sum(m.A for m in self.mechs)
        self.kernel = ctx.declare(wrap_prefix + '/lstm_cell/kernel', (in_dim + self.At + H, 4 * H), 'lstm_kernel')
        self.bias = ctx.declare(wrap_prefix + '/lstm_cell/bias', (4 * H,), 'zeros')
        # flag of the last mechanism created; no mechanisms (enable_attention=False: the bare cell under BasicDecoder,
        # decoder_unimodal.py:319-327 - `wrap_prefix` is then the decoder scope itself): the cell output
        self.output_attention = self.mechs[-1].output_attention if self.mechs else False
        self.out_dim = self.At if self.output_attention else H

    def prepare_memories(self, memories: Sequence[Tuple[torch.Tensor, torch.Tensor]]):
        """keys = memory_layer(values), once per batch (TF computes them at construction)."""
        ctx = self.ctx
        bufs = []
        for md, mem in zip(self.mechs, memories):
            values, mem_len = mem[0], mem[1]
            values_op = mem[2] if len(mem) > 2 and mem[2] is not None else values
            Tm, B, Dm = values.shape
            keys = ops.empty(Tm, B, md.A)
            ops.gemm(values_op.reshape(Tm * B, Dm), ctx.w(md.Wm), keys.view(Tm * B, md.A))
            v = ctx.p(md.v) if md.v else None
            if md.kind == 'normed_bahdanau':
                veff = ops.empty(md.A)
                ops.normed_v_fwd(ctx.p(md.v), ctx.p(md.g), veff)
                v = veff
            mb = ops.MechBuffers(md.kind, values, keys, mem_len, ctx.w(md.Wl),
                                 Wq=ctx.w(md.Wq) if md.Wq else None, v=v,
                                 g=ctx.p(md.g) if md.kind == 'scaled_luong' else None,
                                 bias=ctx.p(md.b) if md.b else None)
            mb.values_op = values_op
            bufs.append(mb)
        return bufs

    def forward(self, x, lens, memories=None, init=None, mech_bufs=None):
        """x [T,B,Dx]; memories [(values [Tm,B,Dm], mem_len)]; init = (c0, h0) or None."""
        T, B, Dx = x.shape
        H = self.H
        ctx = self.ctx
        W, b = ctx.w(self.kernel), ctx.p(self.bias)
        if self.drop is not None:
            x = self.drop.drop_input(x)
        gates = ops.empty(T, B, 4 * H)
        ops.gemm(x.reshape(T * B, Dx), W[:Dx], gates.view(T * B, 4 * H), bias=b)
        self.x = x
        self.bufs = mech_bufs if mech_bufs is not None else self.prepare_memories(memories)
        c0, h0 = init if init is not None else (None, None)
        self.rnn = ops.RnnSeq(T, B, H, lens, gates, W[Dx:], self.bufs, self.output_attention, c0=c0, h0=h0,
                              drop=self.drop)
        out = self.rnn.forward()
        self.final = (self.rnn.cT, self.rnn.hT)
        # operand view of the outputs: the attention vectors are already tf32-rounded (and masked) when they
        # are the output; otherwise the rounded h columns of the state rows (with dropout the state rows hold the
        # state-dropped h, not the emitted one: the consumer rounds the outputs itself)
        self.operand = out if (self.output_attention or self.drop is not None) else self.rnn.S[1:, :, self.At:]
        return out

    def forward_sampled(self, table, true_ids, lens, memories, init, Wd, bd, ss_stream, ss_thr, used_ids, sample_ids):
        """Whole-sequence forward with ScheduledEmbeddingTrainingHelper INSIDE the recurrent kernel (AvsrSampling in
        include/avsr_b200.h): true_ids [T,B] are the teacher-forced decoder inputs, used_ids (pre-filled with them) /
        sample_ids (pre-filled with -1) receive what the helper drew.  Returns the outputs, or None when this layer's
        persistent kernel cannot do it (the caller then advances step by step: begin_stepwise / stepwise_step)."""
        T, B = true_ids.shape
        H, Dx, ctx = self.H, self.Dx, self.ctx
        W, b = ctx.w(self.kernel), ctx.p(self.bias)
        gates = ops.empty(T, B, 4 * H)
        bufs = self.prepare_memories(memories)
        c0, h0 = init if init is not None else (None, None)
        rnn = ops.RnnSeq(T, B, H, lens, gates, W[Dx:], bufs, self.output_attention, c0=c0, h0=h0, drop=self.drop)
        rnn.rng = ctx.rng
        x = ops.empty(T, B, Dx)
        rnn.sampling = ops.Sampling(Wd, bd, table, W[:Dx], b, used_ids, sample_ids, x, ss_stream, ss_thr)
        if not rnn.sampling_fused():
            return None
        ops.embedding_fwd(table, true_ids.reshape(-1), x)
        if self.drop is not None and self.drop.thr_in:
            ops.dropout(x, self.drop.rng, self.drop.stream + 3, self.drop.thr_in, round_out=True, out=x)
        elif ops.tensor_cores_enabled():
            ops.round_tf32(x, x)
        ops.gemm(x.view(T * B, Dx), W[:Dx], gates.view(T * B, 4 * H), bias=b)
        self.x, self.bufs, self.rnn = x, bufs, rnn
        out = rnn.forward()  # rows of x whose input was drawn are replaced by the kernel (operand of dWx in backward)
        self.final = (rnn.cT, rnn.hT)
        self.operand = out if (self.output_attention or self.drop is not None) else rnn.S[1:, :, self.At:]
        return out

    def begin_stepwise(self, T, B, lens, memories, init=None):
        """Scheduled sampling (decoder_unimodal.py:304-309): the decoder inputs are only known step by step.  Sets up
        the whole-sequence buffers; `stepwise_step(t, x_t)` then advances one step and returns that step's output."""
        H, Dx, ctx = self.H, self.Dx, self.ctx
        self.x = ops.empty(T, B, Dx)
        self._gates = ops.empty(T, B, 4 * H)
        self.bufs = self.prepare_memories(memories)
        c0, h0 = init if init is not None else (None, None)
        self.rnn = ops.RnnSeq(T, B, H, lens, self._gates, ctx.w(self.kernel)[Dx:], self.bufs, self.output_attention,
                              c0=c0, h0=h0, drop=self.drop)

    def stepwise_step(self, t, x_t):
        """x_t [B,Dx] un-dropped decoder input of step t."""
        W, b = self.ctx.w(self.kernel), self.ctx.p(self.bias)
        B, Dx, H = x_t.shape[0], self.Dx, self.H
        xt = self.x[t]
        if self.drop is not None and self.drop.thr_in:
            # element index of the whole-sequence mask: (t*B + b)*Dx + column -> offset the flat index by t*B*Dx
            ops.dropout(x_t, self.drop.rng, self.drop.stream + 3, self.drop.thr_in, round_out=True, out=xt,
                        first=t * B * Dx)
        elif ops.tensor_cores_enabled():
            ops.round_tf32(x_t, xt)
        else:
            xt.copy_(x_t)
        ops.gemm(xt, W[:Dx], self._gates[t].view(B, 4 * H), bias=b)
        self.rnn.forward_range(t, t + 1)
        return self.rnn.out[t]

    def end_stepwise(self):
        self.final = (self.rnn.cT, self.rnn.hT)
        out = self.rnn.out
        self.operand = out if (self.output_attention or self.drop is not None) else self.rnn.S[1:, :, self.At:]
        return out

    def step(self, x1, active, mech_bufs, state):
        """One decode step (inference).  x1 [1,B,Dx]; active [B] int32 (1 = run, 0 = carry state,
        emit zeros: impute_finished); state = (c [B,H], S [B,At+H]).  Returns (out [B,O], new state)."""
        _, B, Dx = x1.shape
        H = self.H
        W, b = self.ctx.w(self.kernel), self.ctx.p(self.bias)
        gates = ops.empty(1, B, 4 * H)
        ops.gemm(x1.view(B, Dx), W[:Dx], gates.view(B, 4 * H), bias=b)
        c, S = state
        rnn = ops.RnnSeq(1, B, H, active, gates, W[Dx:], mech_bufs, self.output_attention, c0=c, s0=S)
        out = rnn.forward()
        self.last_step = rnn
        return out[0], (rnn.cT, rnn.S[1])

    def initial_state(self, B, init=None):
        S = ops.zeros(B, self.At + self.H)
        if init is not None:
            S[:, self.At:].copy_(init[1])
            c = init[0].clone()
        else:
            c = ops.zeros(B, self.H)
        return (c, S)

    def backward(self, dout, dstate=None, need_dx=True, want_init_grad=True):
        """Returns (dx, [dmemory per mechanism], (dc0, dh0))."""
        ctx = self.ctx
        T, B, Dx = self.x.shape
        H = self.H
        W = ctx.w(self.kernel)
        gW, gb = ctx.g(self.kernel), ctx.g(self.bias)
        dveff = {}
        for k, (md, mb) in enumerate(zip(self.mechs, self.bufs)):
            mb.dkeys = ops.zeros(mb.Tm, B, mb.A)
            mb.dvalues = ops.zeros(mb.Tm, B, mb.Dm)
            mb.dWl = ctx.g(md.Wl)
            if md.Wq:
                mb.dWq = ctx.g(md.Wq)
                if md.kind == 'normed_bahdanau':
                    dveff[k] = ops.zeros(mb.A)
                    mb.dv = dveff[k]
                    mb.dbias = ctx.g(md.b)
                else:
                    mb.dv = ctx.g(md.v)
            if md.kind == 'scaled_luong':
                mb.dg = ctx.g(md.g)
        dcT, dhT = dstate if dstate is not None else (None, None)
        self.rnn.grad_scale = ctx.grad_scale
        dZ = self.rnn.backward(dout, gW[Dx:], dcT=dcT, dhT=dhT, want_init_grad=want_init_grad, dbias=gb)
        dZ2 = dZ.view(T * B, 4 * H)
        ops.gemm(self.x.reshape(T * B, Dx), dZ2, gW[:Dx], ta=True, beta=1.0)
        dx = None
        if need_dx:
            dx = ops.empty(T, B, Dx)
            ops.gemm(dZ2, W[:Dx], dx.view(T * B, Dx), tb=True)
            if self.drop is not None:
                self.drop.drop_input_grad(dx)
        dmem = []
        for k, (md, mb) in enumerate(zip(self.mechs, self.bufs)):
            v2 = mb.values_op.reshape(mb.Tm * B, mb.Dm)
            dk2 = mb.dkeys.view(mb.Tm * B, mb.A)
            ops.gemm(v2, dk2, ctx.g(md.Wm), ta=True, beta=1.0)  # dWm
            ops.gemm(dk2, ctx.w(md.Wm), mb.dvalues.view(mb.Tm * B, mb.Dm), tb=True, beta=1.0)
            dmem.append(mb.dvalues)
            if k in dveff:
                ops.normed_v_bwd(ctx.p(md.v), ctx.p(md.g), dveff[k], ctx.g(md.v), ctx.g(md.g))
        dinit = (self.rnn.dc0, self.rnn.dh0)
        self.rnn = None
        return dx, dmem, dinit
