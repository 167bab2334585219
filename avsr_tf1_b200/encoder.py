"""Encoders - drop-in for reference avsr/encoder.py (Seq2SeqEncoder :14-196,
AttentiveEncoder :199-335), computing on B200 through libavsr_b200.so."""
from __future__ import annotations

import collections

import numpy as np

import torch

from . import ops
from .attention import add_attention
from .cells import build_rnn_layers
from .layers import BatchNormInput, BuildContext, InstanceNormInput, LSTMLayerOp


class EncoderData(collections.namedtuple("EncoderData", ("outputs", "final_state", "outputs_operand"))):
    """outputs / final_state as in the reference (encoder.py:10).  outputs_operand: the same sequence in the
    form later matrix products read it (tf32-rounded in tensor-core mode; rows past the length unspecified)."""

    def __new__(cls, outputs, final_state, outputs_operand=None):
        return super(EncoderData, cls).__new__(cls, outputs, final_state,
                                               outputs if outputs_operand is None else outputs_operand)


def maybe_list(obj):
    return obj if type(obj) in (list, tuple) else [obj, ]


def _layer_prefixes(scope, n_layers, sub=''):
    pre = f'{scope}/Encoder/' + (f'{sub}/' if sub else '')
    if n_layers == 1:
        return [pre + 'lstm_cell']
    return [pre + f'multi_rnn_cell/cell_{k}/lstm_cell' for k in range(n_layers)]


class InputDenseStack(object):
    """_maybe_add_dense_layers (encoder.py:148-171): Dense(units, activation=tf.nn.selu, use_bias=False,
    kernel_initializer=variance_scaling_initializer(), kernel_regularizer=l2(1e-4)) for every entry of
    `input_dense_layers`, between the input normalisation and the first recurrent layer.  tf.layers names them dense,
    dense_1, ... inside `<scope>/Encoder` (they are called at encoder.py:64, before the state projections of a
    bidirectional encoder get their names)."""
    L2_SCALE = 1e-4

    def __init__(self, ctx: BuildContext, scope, in_dim, units):
        self.ctx = ctx
        self.kernels, d = [], int(in_dim)
        for i, u in enumerate(units):
            name = f'{scope}/Encoder/dense' + ('' if i == 0 else f'_{i}') + '/kernel'
            self.kernels.append(ctx.declare(name, (d, int(u)), 'lstm_kernel'))  # (variance_scaling_initializer(), as the cells)
            d = int(u)
        self.out_dim = d

    def forward(self, x):
        """x [T,B,F] product operand -> [T,B,units[-1]] product operand."""
        ctx = self.ctx
        self._saved = []
        for k in self.kernels:
            T, B, F = x.shape
            W = ctx.w(k)
            y = ops.empty(T, B, W.shape[1])
            ops.gemm(x.reshape(T * B, F), W, y.view(T * B, -1))
            ops.selu_fwd(y, out=y)
            self._saved.append((x, y))
            x = ops.round_tf32(y) if ops.tensor_cores_enabled() else y
        return x

    def backward(self, d):
        """d: gradient wrt the stack's output; accumulates the kernel gradients, returns the gradient wrt its input."""
        ctx = self.ctx
        for k, (x, y) in zip(reversed(self.kernels), reversed(self._saved)):
            T, B, F = x.shape
            dz = ops.selu_bwd(y, d.contiguous())
            if ops.tensor_cores_enabled():
                ops.round_tf32(dz, dz)
            dz2 = dz.view(T * B, -1)
            ops.gemm(x.reshape(T * B, F), dz2, ctx.g(k), ta=True, beta=1.0)
            d = ops.empty(T, B, F)
            ops.gemm(dz2, ctx.w(k), d.view(T * B, F), tb=True)
        self._saved = None
        return d

    def add_l2(self, loss_sumsq, unit_scale):
        """kernel_regularizer: gradient += 1e-4 w; loss_sumsq[0] += (1e-4 / unit_scale) sum w^2 (the caller's slot is later
        multiplied by unit_scale / 2).  Only called when the reference adds REGULARIZATION_LOSSES (seq2seq.py:180-184)."""
        ctx = self.ctx
        for k in self.kernels:
            ops.axpy(self.L2_SCALE, ctx.p(k).reshape(-1), ctx.g(k).reshape(-1))
            tmp = ops.zeros(1)
            ops.sumsq(ctx.p(k).reshape(-1), tmp)
            ops.axpy(self.L2_SCALE / unit_scale, tmp, loss_sumsq)


class Seq2SeqEncoder(object):
    """Input BatchNorm -> stacked uni/bi-directional LSTM (encoder.py:14-196).

    outputs [T,B,H] (Bi: [T,B,2H]); final_state = (c, h) of the LAST layer (Bi: the
    Dense projections of encoder.py:124-141) - the only state the decoders read
    (decoder_unimodal.py:144, decoder_bimodal.py:144-152)."""

    def __init__(self, data, mode, hparams, num_units_per_layer, dropout_probability, ctx: BuildContext = None,
                 scope='', feature_dim=None, **kwargs):
        self._mode, self._hparams = mode, hparams
        self._num_units_per_layer = tuple(num_units_per_layer)
        self._scope, self._ctx = scope, ctx
        # raw lip crops [B,T,h,w,c] enter as flat features (video_processing='features' on a video record)
        self._F = int(feature_dim if feature_dim is not None else np.prod(tuple(data.inputs.shape[2:])))
        # Action-Unit regression head (encoder.py:28-29, 173-189): train mode only
        self._regress_aus = bool(kwargs.get('regress_aus', False)) and mode == 'train'
        self._bn = BatchNormInput(ctx, scope, self._F) if hparams.batch_normalisation is True else None
        self._inorm = InstanceNormInput(ctx, scope, self._F) if hparams.instance_normalisation is True else None
        self._dense = None  # _maybe_add_dense_layers (encoder.py:148-171); default (0,) = identity (avsr.py:38)
        self._rnn_in = self._F
        if hparams.input_dense_layers[0] > 0:
            self._dense = InputDenseStack(ctx, scope, self._F, hparams.input_dense_layers)
            self._rnn_in = self._dense.out_dim
        self.input_gradient = False  # True: keep what the gradient wrt the raw features needs (backward(need_dx=True))
        self._init_encoder()
        if self._regress_aus:
            self._au_W = ctx.declare(f'{scope}/dense/kernel', (self.output_dim, 2), 'glorot')
            self._au_b = ctx.declare(f'{scope}/dense/bias', (2,), 'zeros')
        self._au_dz = None

    def _init_encoder(self):
        hp, ctx, scope = self._hparams, self._ctx, self._scope
        units = self._num_units_per_layer
        L = len(units)
        self._units = units

        def stack(sub):
            cells = maybe_list(build_rnn_layers(
                cell_type=hp.cell_type, num_units_per_layer=units, use_dropout=hp.use_dropout,
                dropout_probability=self._dropout_probability_for(scope), mode=self._mode,
                residual_connections=hp.residual_encoder if not sub else False,
                highway_connections=hp.highway_encoder if not sub else False,
                weight_sharing=hp.encoder_weight_sharing if not sub else False))
            ops_, in_dim = [], self._rnn_in
            for prefix, cell in zip(_layer_prefixes(scope, L, sub), cells):
                share = ops_[cell.share_with] if getattr(cell, 'share_with', None) is not None else None
                op = LSTMLayerOp(ctx, prefix, in_dim, cell.num_units, drop=ctx.drop_state(cell, prefix), share=share)
                op.residual = bool(getattr(cell, 'residual', False))
                op.highway = None
                if op.residual and in_dim != cell.num_units:
                    raise ValueError('residual connections need layers of equal width (ResidualWrapper adds input and output)')
                if getattr(cell, 'highway', False):
                    # tf.contrib.rnn.HighwayWrapper (cells.py:89-90): `carry_w` [in, in] (default glorot initialiser) and
                    # `carry_b` (constant 1) live in the position's own scope (`.../cell_k/`), also when the cell is shared
                    if in_dim != cell.num_units:
                        raise ValueError('highway connections need layers of equal width (HighwayWrapper mixes input and output)')
                    scope_k = prefix.rsplit('/', 1)[0]
                    op.highway = (ctx.declare(scope_k + '/carry_w', (in_dim, in_dim), 'glorot'),
                                  ctx.declare(scope_k + '/carry_b', (in_dim,), 'const:1.0'))
                ops_.append(op)
                in_dim = cell.num_units
            return ops_

        if hp.encoder_type == 'unidirectional':
            self._fw, self._bw = stack(''), None
            self.output_dim = units[-1]
        elif hp.encoder_type == 'bidirectional':
            if hp.cell_type != 'lstm':
                raise ValueError('BiRNN fusion strategy not implemented for this cell')
            if L == 1:
                raise ValueError('the reference cannot build a 1-layer bidirectional LSTM encoder '
                                 '(encoder.py:125-133 indexes the state as a tuple of layers)')
            self._fw, self._bw = stack('fw'), stack('bw')
            dec = hp.decoder_units_per_layer[0]
            self._proj = []
            nd = len(self._dense.kernels) if self._dense is not None else 0  # the input dense stack is named first
            for i in range(nd, nd + 2 * L):  # encoder.py:124-141: dense, dense_1, ... (c then h per layer)
                name = f'{scope}/Encoder/dense' + ('' if i == 0 else f'_{i}') + '/kernel'
                self._proj.append(ctx.declare(name, (2 * units[(i - nd) // 2], dec), 'glorot'))
            self.output_dim = 2 * units[-1]
        else:
            raise Exception('Allowed encoder types: `unidirectional`, `bidirectional`')

    def _dropout_probability_for(self, scope):
        hp = self._hparams
        return hp.video_encoder_dropout_probability if scope == 'video' else hp.audio_encoder_dropout_probability

    # ---- compute -------------------------------------------------------------
    def _normalised_inputs(self, inputs, batch_major):
        """Input BN (or the plain operand rounding) -> frame-major [T,B,F] product operand."""
        train = self._mode == 'train'
        if self._bn is not None:
            # tf32-rounded in tensor-core mode; xhat is only stored if the gradient wrt the raw features is wanted
            # (or if layer 0 drops its input: then dgamma / dbeta need the gradient wrt the normalised features)
            x = self._bn.forward(inputs, train, batch_major=batch_major,
                                 keep_xhat=self.input_gradient or self._layer0_drops_input())
        else:
            if batch_major:
                inputs = ops.transpose01(inputs)
            x = inputs if self._inorm is not None else (ops.round_tf32(inputs) if ops.tensor_cores_enabled() else inputs)
        if self._inorm is not None:  # instance_norm on the (batch-normalised) features (encoder.py:51-55)
            x = self._inorm.forward(x)
        return self._dense.forward(x) if self._dense is not None else x

    def _layer0_ops(self):
        first = [self._fw[0]] if self._fw else [self._top]  # (an AV-Align encoder of one layer: the attention cell)
        return first + ([self._bw[0]] if self._bw is not None else [])

    def _layer0_drops_input(self):
        """True when the gradient wrt the normalised features has to be formed explicitly (dgamma / dbeta cannot be read
        off the layer-0 weight gradient): layer 0 drops its input, or a dense stack / the instance normalisation sits between
        the two."""
        if self._mode == 'train' and (self._dense is not None or self._inorm is not None or
                                      getattr(self, 'explicit_bn_backward', False)):
            return True  # (explicit_bn_backward: set by Seq2SeqModel._guard_bn_shortcut when a gamma came close to zero)
        return self._mode == 'train' and any(op.drop is not None and op.drop.thr_in for op in self._layer0_ops())

    def _round_outputs(self, out, op):
        """Operand view of a stack's outputs: with a DropoutWrapper the layer hands back its exact outputs."""
        if op.drop is not None and ops.tensor_cores_enabled():
            return ops.round_tf32(out)
        return op.operand

    def forward(self, inputs, inputs_len, batch_major=False):
        """inputs [T,B,F] frame-major (or [B,T,F] with batch_major=True); returns EncoderData."""
        ctx = self._ctx
        self._lens = inputs_len
        x = self._normalised_inputs(inputs, batch_major)

        def backward_stack():
            curb, curb_op = None, ops.reverse_sequence(x, inputs_len)
            for op in self._bw:
                curb = op.forward(curb_op, inputs_len)
                curb_op = op.operand
            return ops.reverse_sequence(curb, inputs_len)
        # bidirectional_dynamic_rnn (encoder.py:110) = two independent deep stacks: side by side when they fit the GPU
        side = ctx.fork(1) if (self._bw is not None and ctx.parallel_chains) else None
        if side is not None:
            with torch.cuda.stream(side):
                outb = backward_stack()
        cur, cur_op = None, x
        for op in self._fw:
            out = op.forward(cur_op, inputs_len)
            if getattr(op, 'highway', None):  # HighwayWrapper (cells.py:89-90) around the (dropout-wrapped) cell
                T_, B_, D_ = cur.shape
                pre = ops.empty(T_, B_, D_)
                # (the carry product's operand: rounded here - under dropout a layer hands back its exact outputs)
                x_op = ops.round_tf32(cur) if ops.tensor_cores_enabled() else cur
                ops.gemm(x_op.reshape(T_ * B_, D_), ctx.w(op.highway[0]), pre.view(T_ * B_, D_), bias=ctx.p(op.highway[1]))
                op.hw_saved = (cur, x_op, pre, out)
                cur, cur_op = ops.highway_fwd(cur, pre, out)
            elif getattr(op, 'residual', False):  # ResidualWrapper (cells.py:91-92): + the layer's (un-dropped) input
                cur = out + cur
                cur_op = ops.round_tf32(cur) if ops.tensor_cores_enabled() else cur
            else:
                cur, cur_op = out, op.operand
        if side is not None:
            ctx.join(1)
        elif self._bw is not None:
            outb = backward_stack()
        if self._bw is None:
            self._outputs = cur
            wrapped = getattr(self._fw[-1], 'residual', False) or getattr(self._fw[-1], 'highway', None)
            self._outputs_op = cur_op if wrapped else self._round_outputs(cur, self._fw[-1])
            self._final = self._fw[-1].final
        else:
            T, B, H = cur.shape
            out = ops.empty(T, B, 2 * H)
            out[:, :, :H].copy_(cur)
            out[:, :, H:].copy_(outb)
            self._outputs = out
            self._outputs_op = ops.round_tf32(out) if ops.tensor_cores_enabled() else out
            cf, hf = self._fw[-1].final
            cb, hb = self._bw[-1].final
            self._cat_c = ops.empty(B, 2 * H)
            self._cat_h = ops.empty(B, 2 * H)
            self._cat_c[:, :H].copy_(cf); self._cat_c[:, H:].copy_(cb)
            self._cat_h[:, :H].copy_(hf); self._cat_h[:, H:].copy_(hb)
            dec = self._hparams.decoder_units_per_layer[0]
            pc, ph = ops.empty(B, dec), ops.empty(B, dec)
            ops.gemm(self._cat_c, ctx.w(self._proj[-2]), pc)
            ops.gemm(self._cat_h, ctx.w(self._proj[-1]), ph)
            self._final = (pc, ph)
        return self.get_data()

    def get_data(self):
        return EncoderData(outputs=self._outputs, final_state=self._final, outputs_operand=self._outputs_op)

    # ---- Action-Unit regression head (encoder.py:173-189) ---------------------------------------
    def au_loss_forward(self, aus, scale_dev, loss_sum):
        """aus [B,T,2] (payload, batch-major); scale_dev: device scalar au_loss_weight / non-zero weight count;
        loss_sum[0] += masked sum of squared errors.  Keeps d(loss)/d(pre-sigmoid) for au_loss_backward."""
        ctx = self._ctx
        T, B, D = self._outputs.shape
        z = ops.empty(T, B, 2)
        ops.gemm(self._outputs.view(T * B, D), ctx.p(self._au_W), z.view(T * B, 2), bias=ctx.p(self._au_b))
        self._au_dz = ops.empty(T, B, 2)
        ops.au_loss(z, aus, self._lens, scale_dev, loss_sum, self._au_dz)

    def au_loss_backward(self, doutputs):
        """Adds the head's gradient to doutputs (allocated when None); accumulates the head's weight gradients."""
        ctx = self._ctx
        T, B, D = self._outputs.shape
        dz2 = self._au_dz.view(T * B, 2)
        ops.gemm(self._outputs.view(T * B, D), dz2, ctx.g(self._au_W), ta=True, beta=1.0)
        ops.colsum(dz2, ctx.g(self._au_b))
        if doutputs is None:
            doutputs = ops.zeros(T, B, D)
        ops.gemm(dz2, ctx.p(self._au_W), doutputs.view(T * B, D), tb=True, beta=1.0)
        self._au_dz = None
        return doutputs

    def backward(self, doutputs, dfinal_state=None, need_dx=False):
        """doutputs [T,B,out_dim] or None; dfinal_state = (dc, dh) wrt final_state or None."""
        ctx = self._ctx
        if self._bw is None:
            d = doutputs if doutputs is not None else ops.zeros(*self._outputs.shape)
            n = len(self._fw)
            need_dx = need_dx or self._layer0_drops_input()
            for i in range(n - 1, -1, -1):
                need = (i > 0) or need_dx
                op = self._fw[i]
                if getattr(op, 'highway', None):
                    x_in, x_op, pre, out = op.hw_saved
                    op.hw_saved = None
                    T_, B_, D_ = x_in.shape
                    dx_carry, dcell, dpre = ops.highway_bwd(d, x_in, pre, out)
                    dpre2 = dpre.view(T_ * B_, D_)
                    ops.gemm(x_op.reshape(T_ * B_, D_), dpre2, ctx.g(op.highway[0]), ta=True, beta=1.0)
                    ops.colsum(dpre2, ctx.g(op.highway[1]))
                    dn = op.backward(dcell, dfinal_state if i == n - 1 else None, need_dx=True)
                    ops.gemm(dpre2, ctx.w(op.highway[0]), dn.view(T_ * B_, D_), tb=True, beta=1.0)
                    ops.axpy(1.0, dx_carry, dn)
                    d = dn
                    continue
                dn = op.backward(d, dfinal_state if i == n - 1 else None, need_dx=need)
                if getattr(op, 'residual', False) and dn is not None:
                    ops.axpy(1.0, d, dn)  # the shortcut around the cell
                d = dn
            dx = d
        else:
            T, B, H2 = self._outputs.shape
            H = H2 // 2
            if doutputs is None:
                doutputs = ops.zeros(T, B, H2)
            dsf = dsb = None
            if dfinal_state is not None:
                dpc, dph = dfinal_state
                ops.gemm(self._cat_c, dpc, ctx.g(self._proj[-2]), ta=True, beta=1.0)
                ops.gemm(self._cat_h, dph, ctx.g(self._proj[-1]), ta=True, beta=1.0)
                dcc, dch = ops.empty(B, H2), ops.empty(B, H2)
                ops.gemm(dpc, ctx.w(self._proj[-2]), dcc, tb=True)
                ops.gemm(dph, ctx.w(self._proj[-1]), dch, tb=True)
                dsf = (dcc[:, :H].contiguous(), dch[:, :H].contiguous())
                dsb = (dcc[:, H:].contiguous(), dch[:, H:].contiguous())
            df = doutputs[:, :, :H].contiguous()
            db = ops.reverse_sequence(doutputs[:, :, H:].contiguous(), self._lens)
            n = len(self._fw)
            need_dx = need_dx or self._layer0_drops_input()

            def stack_backward(stack, d, ds):
                for i in range(n - 1, -1, -1):
                    d = stack[i].backward(d, ds if i == n - 1 else None, need_dx=(i > 0) or need_dx)
                return d
            if ctx.parallel_chains:  # the two stacks are independent in the backward pass too
                # `db` was allocated on THIS stream and is read by kernels of the side stream: it must stay referenced
                # until the join, or the caching allocator hands its block to the forward stack's allocations while
                # the side kernels still read it (the allocator only orders reuse within the allocating stream)
                keep = db
                with torch.cuda.stream(ctx.fork(1)):
                    db_new = stack_backward(self._bw, db, dsb)
                df = stack_backward(self._fw, df, dsf)
                ctx.join(1)
                db = db_new
                del keep
            else:
                df = stack_backward(self._fw, df, dsf)
                db = stack_backward(self._bw, db, dsb)
            dx = None
            if need_dx:
                dxb = ops.reverse_sequence(db, self._lens)
                ops.axpy(1.0, dxb, df)
                dx = df
        return self._input_bn_backward(dx, need_dx)

    def _input_bn_backward(self, dx, need_dx):
        """Gradients of the input normalisation.  Its output only feeds the layer-0 gate product(s), so dgamma / dbeta
        follow from the layer-0 weight gradients (BatchNormInput.backward_from_layer0) and the [T*B,F] gradient wrt the
        normalised features is only formed when the caller asks for the gradient wrt the raw features."""
        if self._dense is not None and dx is not None:
            dx = self._dense.backward(dx)
        if self._inorm is not None and dx is not None:
            dx = self._inorm.backward(dx)
        if self._bn is None:
            return dx
        if self._layer0_drops_input() and not self.input_gradient:
            self._bn.backward(dx, need_dx=False)  # dx: gradient wrt the normalised features (input dropout undone)
            return None
        if need_dx:
            if not self.input_gradient:
                raise Exception('encoder: set input_gradient = True before forward to get the gradient wrt the features')
            return self._bn.backward(dx, need_dx=True)
        self._bn.backward_from_layer0(self._layer0_ops())
        return None


class AttentiveEncoder(Seq2SeqEncoder):
    """AV-Align (encoder.py:199-335, https://arxiv.org/abs/1809.01728): the TOP audio layer is
    an AttentionWrapper attending to the video encoder outputs (audio attends to video)."""

    def __init__(self, data, mode, hparams, num_units_per_layer, attended_memory_depth, dropout_probability,
                 ctx: BuildContext = None, scope='audio', feature_dim=None):
        self._attended_memory_depth = int(attended_memory_depth)
        super(AttentiveEncoder, self).__init__(data, mode, hparams, num_units_per_layer, dropout_probability,
                                               ctx=ctx, scope=scope, feature_dim=feature_dim)

    def _init_encoder(self):
        hp, ctx, scope = self._hparams, self._ctx, self._scope
        units = self._num_units_per_layer
        L = len(units)
        if hp.encoder_type != 'unidirectional':
            raise Exception('AttentiveEncoder: only `unidirectional` is implemented (encoder.py:229)')
        cells = maybe_list(build_rnn_layers(
            cell_type=hp.cell_type, num_units_per_layer=units, use_dropout=hp.use_dropout,
            dropout_probability=hp.audio_encoder_dropout_probability, mode=self._mode, as_list=True))
        self._fw, in_dim = [], self._rnn_in
        for k in range(L - 1):
            prefix = f'{scope}/Encoder/multi_rnn_cell/cell_{k}/lstm_cell'
            self._fw.append(LSTMLayerOp(ctx, prefix, in_dim, cells[k].num_units, drop=ctx.drop_state(cells[k], prefix)))
            in_dim = cells[k].num_units
        wrap = f'{scope}/Encoder/multi_rnn_cell/cell_{L - 1}/attention_wrapper' if L > 1 \
            else f'{scope}/Encoder/attention_wrapper'
        self._top = add_attention(cells[-1], attention_types=hp.attention_type[0], num_units=units[-1],
                                  memory_depths=[self._attended_memory_depth], ctx=ctx, wrap_prefix=wrap,
                                  mem_layer_names=[f'{scope}/Encoder/memory_layer/kernel'], in_dim=in_dim,
                                  fusion_type='linear_fusion')
        self._bw = None
        self.output_dim = self._top.out_dim

    def forward_lower(self, inputs, inputs_len, batch_major=False):
        """Input BN + the plain layers under the attention layer (independent of the video stream, so the
        model can run it concurrently with the video encoder)."""
        self._lens = inputs_len
        x = self._normalised_inputs(inputs, batch_major)
        cur_op = x
        for op in self._fw:
            op.forward(cur_op, inputs_len)
            cur_op = op.operand
        self._lower_out_op = cur_op

    def forward_top(self, attended_memory, attended_memory_length, attended_memory_operand=None):
        self._outputs = self._top.forward(
            self._lower_out_op, self._lens,
            memories=[(attended_memory, attended_memory_length, attended_memory_operand)])
        self._outputs_op = self._round_outputs(self._outputs, self._top) if not self._top.output_attention \
            else self._top.operand
        self._final = self._top.final  # wrapper stripped: cell state only (encoder.py:314-330)
        self.alignment_history = self._top.bufs[0].align    # [T_audio, B, T_video] device tensor
        self.attention_contexts = self._top.bufs[0].hc      # [T_audio, B, H + Dm]
        if self._hparams.write_attention_alignment:
            # encoder.py:296-310: alignment_history stacked and transposed to [B, T_video, T_audio, 1]; the summary is
            # the image 1 - alignment (host arrays here instead of tf.Summary protos)
            self.attention_alignment = self.alignment_history.permute(1, 2, 0).unsqueeze(-1).cpu().numpy()
            self.attention_summary = 1.0 - self.attention_alignment
        return self.get_data()

    def forward(self, inputs, inputs_len, attended_memory=None, attended_memory_length=None,
                attended_memory_operand=None):
        self.forward_lower(inputs, inputs_len)
        return self.forward_top(attended_memory, attended_memory_length, attended_memory_operand)

    def backward_top(self, doutputs, dfinal_state=None):
        """Returns (gradient wrt the lower layers' output, dmemory = gradient wrt the video encoder outputs)."""
        if doutputs is None:
            doutputs = ops.zeros(*self._outputs.shape)
        d, dmem, _ = self._top.backward(doutputs, dfinal_state, need_dx=True, want_init_grad=False)
        return d, dmem[0]

    def backward_lower(self, d, need_dx=False):
        need_dx = need_dx or self._layer0_drops_input()
        for i in range(len(self._fw) - 1, -1, -1):
            need = (i > 0) or need_dx
            d = self._fw[i].backward(d, None, need_dx=need)
        return self._input_bn_backward(d if need_dx else None, need_dx)

    def backward(self, doutputs, dfinal_state=None, need_dx=False):
        """Returns (dx, dmemory) - dmemory is the gradient wrt the video encoder outputs."""
        d, dmem = self.backward_top(doutputs, dfinal_state)
        return self.backward_lower(d, need_dx), dmem
