"""Visual front-end - drop-in for reference avsr/video.py (`resnet_cnn` :143-195 applied per frame by `cnn_layers`
:224-248).  NHWC like the reference (`data_format='channels_last'`).

    layer0: conv3x3(C -> f0) -> BN-ReLU
    res_block_0 (skip_bn): conv3x3 -> BN-ReLU -> conv3x3, + input
    res_block_k (k >= 1): BN-ReLU(x) -> conv3x3 stride 2 -> BN-ReLU -> conv3x3, + conv1x1 stride 2 of the RAW x
    flatten: VALID conv over what is left of the image -> cnn_dense_units, ReLU

The narrow convolutions (8 / 16 output channels at 36x36 and 18x18: nearly all pixels) run on direct kernels
(avsr_conv2d_direct, avsr_conv2d_wgrad; exact fp32); the 32- / 64-channel ones are avsr_im2col + avsr_gemm (bias in the
product's epilogue), their gradients avsr_gemm + avsr_colsum + avsr_col2im.  BN is the same stats / apply pair as the
input normalisation (rows = N*H*W; eps 1e-5, momentum 0.98; all-reduced sums under data parallelism).  im2col buffers
are not kept: the backward pass rebuilds them from the saved layer inputs.  That is the exact-fp32 path (parity against the oracle at 1e-3).

In tensor-core mode (the default) every convolution runs on the implicit-GEMM `mma.sync` TF32 kernels of csrc/conv_mma.cu
(forward, both input gradients, weight gradient: no im2col / col2im buffers) and the batch_norm_relu layers never
materialise: a convolution accumulates the statistics of the batch norm that follows in its epilogue, its consumers apply
the BN-ReLU while they stage their input, the ReLU mask and the statistics pass of the BN backward ride in the epilogue of
the convolution that forms the incoming gradient, and the residual `tf.add`s are epilogue adds
(`ResNetCNN._forward_fused / _backward_fused`).

`2dconv_cnn` / `3dconv_cnn` (video.py:108-140, 198-221) are not built (no reference script selects them)."""
from __future__ import annotations

import os

import torch

from . import ops
from .layers import BuildContext

BN_EPS = 1e-5        # video.py:10
BN_MOMENTUM = 0.98   # video.py:10
L2_SCALE = 1e-3      # video.py:27 kernel_regularizer


class _Conv(object):
    def __init__(self, ctx: BuildContext, name, kh, kw, cin, cout, stride=1, padding='SAME'):
        self.ctx, self.kh, self.kw, self.cin, self.cout, self.stride, self.padding = ctx, kh, kw, cin, cout, stride, padding
        self.kernel = ctx.declare(name + '/kernel', (kh, kw, cin, cout), 'conv_kernel')
        self.bias = ctx.declare(name + '/bias', (cout,), 'zeros')

    @property
    def direct(self):
        """8 / 16 output channels (nearly all pixels of the network): direct convolution kernels, no im2col buffer."""
        return (self.cout in (8, 16) and self.kh * self.kw * self.cin * self.cout <= 2304
                and not os.environ.get('AVSR_CNN_IM2COL'))

    @property
    def tensor_core(self):
        """Tensor-core mode: implicit-GEMM kernels of csrc/conv_mma.cu (forward, both input gradients, weight gradient)."""
        return (ops.tensor_cores_enabled() and self.padding == 'SAME' and not os.environ.get('AVSR_CNN_NO_TC')
                and ops.conv2d_tc_supported(self.cin, self.cout, self.kh, self.kw, self.stride))

    def forward(self, x, residual=None, stats=None):
        """x [N,H,W,Cin] -> [N,Ho,Wo,Cout] (exact fp32 out; the operand copy of x is made by im2col).
        residual: added to the output (residual_block's tf.add); stats [2 Cout]: += (sum, sum of squares) of the output
        per channel, for the batch norm that follows - both fused into the tensor-core kernel, separate passes otherwise."""
        ctx = self.ctx
        self._x = x
        if self.tensor_core:
            N, H, W, _ = x.shape
            Ho, Wo, pt, pl = ops.conv_geometry(H, W, self.kh, self.kw, self.stride, self.padding)
            return ops.conv2d_tc(x, ctx.p(self.kernel).view(-1, self.cout), ctx.p(self.bias), self.kh, self.kw, self.stride,
                                 pt, pl, Ho, Wo, residual=residual, stats=stats)
        y = self._forward_plain(x)
        if residual is not None:
            ops.axpy(1.0, residual, y)
        if stats is not None:
            ops.bn_stats(y.view(-1, self.cout), stats)
        return y

    # --- building blocks of the fused tensor-core path (ResNetCNN._forward_fused / _backward_fused) ---
    def conv_tc(self, x, **kw):
        H, W = x.shape[1], x.shape[2]
        Ho, Wo, pt, pl = ops.conv_geometry(H, W, self.kh, self.kw, self.stride, self.padding)
        return ops.conv2d_tc(x, self.ctx.p(self.kernel).view(-1, self.cout), self.ctx.p(self.bias), self.kh, self.kw,
                             self.stride, pt, pl, Ho, Wo, **kw)

    def wgrad_tc(self, x, dy, in_bn=None):
        """kernel and bias gradients from the layer input x (read through in_bn) and dy."""
        ctx = self.ctx
        ops.conv2d_wgrad_tc(x, dy, self.kh, self.kw, self.stride, self.padding, ctx.g(self.kernel).view(-1, self.cout),
                            in_bn=in_bn, dbias=ctx.g(self.bias))

    def dgrad_tc(self, dy, in_hw, **kw):
        """gradient wrt the layer input [N, in_hw, in_hw, cin]: stride-1 convolution of dy (zero-stuffed for a stride-2
        layer) with the kernel flipped in space and transposed in channels, padding k - 1 - pad."""
        H, W = in_hw
        _, _, pt, pl = ops.conv_geometry(H, W, self.kh, self.kw, self.stride, self.padding)
        wt = self.ctx.p(self.kernel).flip(0, 1).permute(0, 1, 3, 2).contiguous().view(-1, self.cin)
        return ops.conv2d_tc(dy, wt, None, self.kh, self.kw, 1, self.kh - 1 - pt, self.kw - 1 - pl, H, W,
                             in_dilation=self.stride, **kw)

    def _forward_plain(self, x):
        ctx = self.ctx
        if self.direct:
            return ops.conv2d_direct(x, ctx.p(self.kernel).view(-1, self.cout), ctx.p(self.bias), self.kh, self.kw,
                                     self.stride, self.padding)
        cols, geom = ops.im2col(x, self.kh, self.kw, self.stride, self.padding)
        N, Ho, Wo = geom[0], geom[9], geom[10]
        y = ops.empty(N, Ho, Wo, self.cout)
        ops.gemm(cols, ctx.w(self.kernel).view(-1, self.cout), y.view(-1, self.cout), bias=ctx.p(self.bias))
        return y

    def _geometry(self):
        N, H, W, C = self._x.shape
        Ho, Wo, pt, pl = ops.conv_geometry(H, W, self.kh, self.kw, self.stride, self.padding)
        return (N, H, W, C, self.kh, self.kw, self.stride, pt, pl, Ho, Wo)

    def backward(self, dy, need_dx=True):
        ctx = self.ctx
        if self.tensor_core:
            dyc = dy.contiguous()
            x = self._x
            self._x = None
            N, H, W, _ = x.shape
            Ho, Wo, pt, pl = ops.conv_geometry(H, W, self.kh, self.kw, self.stride, self.padding)
            ops.conv2d_wgrad_tc(x, dyc, self.kh, self.kw, self.stride, self.padding, ctx.g(self.kernel).view(-1, self.cout))
            ops.colsum(dyc.view(-1, self.cout), ctx.g(self.bias))
            if not need_dx:
                return None
            if ops.conv2d_tc_supported(self.cout, self.cin, self.kh, self.kw, 1):
                # gradient wrt the input = stride-1 convolution of dy (zero-stuffed for a stride-2 layer) with the kernel
                # flipped in space and transposed in channels, padding k - 1 - pad
                wt = ctx.p(self.kernel).flip(0, 1).permute(0, 1, 3, 2).contiguous().view(-1, self.cin)
                return ops.conv2d_tc(dyc, wt, None, self.kh, self.kw, 1, self.kh - 1 - pt, self.kw - 1 - pl, H, W,
                                     in_dilation=self.stride)
            dcols = ops.empty(dyc.numel() // self.cout, self.kh * self.kw * self.cin)
            ops.gemm(dyc.view(-1, self.cout), ctx.p(self.kernel).view(-1, self.cout), dcols, tb=True)
            return ops.col2im(dcols, (N, H, W, self.cin, self.kh, self.kw, self.stride, pt, pl, Ho, Wo))
        if self.direct:
            dyc = dy.contiguous()
            ops.conv2d_wgrad(self._x, dyc, self.kh, self.kw, self.stride, self.padding,
                             ctx.g(self.kernel).view(-1, self.cout))
            ops.colsum(dyc.view(-1, self.cout), ctx.g(self.bias))
            geom = self._geometry()
            self._x = None
            if not need_dx:
                return None
            if self.stride == 1 and self.padding == 'SAME' and self.cin in (8, 16) and self.kh % 2 == 1:
                # gradient wrt the input of a stride-1 SAME convolution = the same convolution of dy with the kernel
                # flipped in space and transposed in channels
                wt = ctx.p(self.kernel).flip(0, 1).permute(0, 1, 3, 2).contiguous().view(-1, self.cin)
                return ops.conv2d_direct(dyc, wt, None, self.kh, self.kw, 1, 'SAME')
            dcols = ops.empty(dyc.numel() // self.cout, self.kh * self.kw * self.cin)
            ops.gemm(dyc.view(-1, self.cout), ctx.p(self.kernel).view(-1, self.cout), dcols, tb=True)
            return ops.col2im(dcols, geom)
        cols, geom = ops.im2col(self._x, self.kh, self.kw, self.stride, self.padding)  # rebuilt, not stored
        d2 = dy.reshape(-1, self.cout)
        if ops.tensor_cores_enabled():
            d2 = ops.round_tf32(d2)  # operand of the two products below
        ops.gemm(cols, d2, ctx.g(self.kernel).view(-1, self.cout), ta=True, beta=1.0)
        ops.colsum(d2, ctx.g(self.bias))
        self._x = None
        if not need_dx:
            return None
        dcols = torch.empty_like(cols)
        ops.gemm(d2, ctx.w(self.kernel).view(-1, self.cout), dcols, tb=True)
        return ops.col2im(dcols, geom)


class _BatchNormRelu(object):
    """batch_norm_relu (video.py:4-15) over NHWC: per-channel statistics over N, H, W."""

    def __init__(self, ctx: BuildContext, name, C):
        self.ctx, self.C = ctx, C
        self.gamma = ctx.declare(name + '/gamma', (C,), 'ones')
        self.beta = ctx.declare(name + '/beta', (C,), 'zeros')
        self.mm = ctx.declare(name + '/moving_mean', (C,), 'zeros', trainable=False)
        self.mv = ctx.declare(name + '/moving_variance', (C,), 'ones', trainable=False)

    # --- fused path: the layer never materialises; the convolutions around it apply these coefficients ---
    def coef(self, stats, rows, train):
        """[4C] (scale, shift, invstd, -mean invstd) from the statistics the producing convolution accumulated (training;
        updates the moving statistics) or from the moving statistics (inference)."""
        ctx = self.ctx
        if not train:
            return ops.bn_coef_eval(ctx.p(self.gamma), ctx.p(self.beta), ctx.p(self.mm), ctx.p(self.mv), BN_EPS)
        count = float(rows)
        if ctx.world_size > 1:
            ctx.allreduce(stats)
            count *= ctx.batch_scale  # rows of the GLOBAL batch (layers.BuildContext.batch_scale)
        self.count = count
        return ops.bn_finalize(stats, count, ctx.p(self.gamma), ctx.p(self.beta), BN_EPS, BN_MOMENTUM, ctx.p(self.mm),
                               ctx.p(self.mv))

    def backward_fused(self, d, u, coef, sums2, residual=None):
        """d: gradient wrt the layer output already masked by its ReLU, sums2 = (sum d, sum d xhat) (both from the epilogue of
        the convolution that produced d); u: the layer's input.  Returns the gradient wrt u (+ residual), in place of d."""
        ctx, C = self.ctx, self.C
        local = sums2
        if ctx.world_size > 1:
            local = sums2.clone()
            ctx.allreduce(sums2)
        du = ops.bn_relu_bwd_apply(d, u, coef, sums2, self.count, residual=residual, out=d)
        ops.axpy(1.0, local[C:], ctx.g(self.gamma))
        ops.axpy(1.0, local[:C], ctx.g(self.beta))
        return du

    def forward(self, x, train):
        ctx, C = self.ctx, self.C
        x2 = x.reshape(-1, C)
        y = torch.empty_like(x)
        if not train:
            ops.bn_apply_eval(x2, ctx.p(self.gamma), ctx.p(self.beta), ctx.p(self.mm), ctx.p(self.mv), BN_EPS,
                              y.view(-1, C))
            return ops.relu_fwd(y, out=y)
        sums = ops.zeros(2 * C)
        ops.bn_stats(x2, sums)
        count = float(x2.shape[0])
        if ctx.world_size > 1:
            ctx.allreduce(sums)
            count *= ctx.batch_scale  # rows of the GLOBAL batch (layers.BuildContext.batch_scale)
        self.xhat, self.invstd, self.count = torch.empty_like(x2), ops.empty(C), count
        ops.bn_apply_train(x2, sums, count, ctx.p(self.gamma), ctx.p(self.beta), BN_EPS, BN_MOMENTUM, y.view(-1, C),
                           self.xhat, self.invstd, ctx.p(self.mm), ctx.p(self.mv))
        self.y = ops.relu_fwd(y, out=y)
        return self.y

    def backward(self, dy):
        ctx, C = self.ctx, self.C
        d = ops.relu_bwd(self.y, dy).view(-1, C)
        sums2 = ops.zeros(2 * C)
        ops.bn_bwd_stats(d, self.xhat, sums2)
        local = sums2
        if ctx.world_size > 1:
            local = sums2.clone()
            ctx.allreduce(sums2)
        dx = torch.empty_like(d)
        ops.bn_bwd_apply(d, self.xhat, sums2, self.count, ctx.p(self.gamma), self.invstd, dx, None, None)
        ops.axpy(1.0, local[C:], ctx.g(self.gamma))
        ops.axpy(1.0, local[:C], ctx.g(self.beta))
        self.xhat = self.y = None
        return dx.view(dy.shape)


class ResNetCNN(object):
    """resnet_cnn(inputs, is_training, cnn_dense_units, cnn_filters) of video.py:143-195 with explicit backward."""

    def __init__(self, ctx: BuildContext, height, width, channels, cnn_filters=(8, 16, 32, 64), cnn_dense_units=128,
                 prefix='CNN/'):
        self.ctx = ctx
        self.in_shape = (int(height), int(width), int(channels))
        self.out_dim = int(cnn_dense_units)
        f = list(cnn_filters)
        self.layer0 = _Conv(ctx, prefix + 'layer0', 3, 3, channels, f[0])
        self.layer0_bn = _BatchNormRelu(ctx, prefix + 'layer0_bn', f[0])
        self.blocks = []
        cin, h, w = f[0], int(height), int(width)
        for k, filt in enumerate(f):
            name = prefix + 'res_block_%d' % k
            blk = {}
            stride = 1 if k == 0 else 2
            if k > 0:  # skip_bn only for block 0; projection shortcut for the strided blocks
                blk['first_bn'] = _BatchNormRelu(ctx, name + '_first_bn', cin)
                blk['shortcut'] = _Conv(ctx, name + '_shortcut', 1, 1, cin, filt, stride=2)
                h, w = -(-h // 2), -(-w // 2)
            blk['conv1'] = _Conv(ctx, name + '_conv1', 3, 3, cin, filt, stride=stride)
            blk['second_bn'] = _BatchNormRelu(ctx, name + '_second_bn', filt)
            blk['conv2'] = _Conv(ctx, name + '_conv2', 3, 3, filt, filt)
            self.blocks.append(blk)
            cin = filt
        self.flatten = _Conv(ctx, prefix + 'flatten', h, w, cin, self.out_dim, padding='VALID')
        self.kernels = [c.kernel for c in self._convs()]

    def _convs(self):
        out = [self.layer0]
        for blk in self.blocks:
            out += [blk[k] for k in ('shortcut', 'conv1', 'conv2') if k in blk]
        return out + [self.flatten]

    def _fused_ok(self):
        """Tensor-core mode and every layer shape covered by csrc/conv_mma.cu: the fused path (no BN / ReLU / add passes)."""
        if os.environ.get('AVSR_CNN_NO_FUSE'):
            return False
        convs = self._convs()[:-1]
        if not all(c.tensor_core for c in convs):
            return False
        if not all(ops.conv2d_tc_supported(c.cout, c.cin, c.kh, c.kw, 1) for c in convs[1:]):  # input gradients
            return False
        return all(1024 % c.cout == 0 for c in convs)

    def _forward_fused(self, frames, train):
        """Every convolution writes its raw output once and accumulates the statistics of the batch norm that follows; the
        BN-ReLU itself is applied by the consumers while they stage their input (csrc/conv_mma.cu)."""
        st = (lambda C: ops.zeros(2 * C)) if train else (lambda C: None)
        rows = lambda t: t.numel() // t.shape[-1]
        sv = []
        blk = self.blocks[0]
        s0 = st(self.layer0.cout)
        y0 = self.layer0.conv_tc(frames, stats=s0)
        c0 = self.layer0_bn.coef(s0, rows(y0), train)
        s1 = st(blk['conv1'].cout)
        y1 = blk['conv1'].conv_tc(y0, in_bn=c0, stats=s1)
        c1 = blk['second_bn'].coef(s1, rows(y1), train)
        sn = st(blk['conv2'].cout) if len(self.blocks) > 1 else None
        o = blk['conv2'].conv_tc(y1, in_bn=c1, residual=y0, res_bn=c0, stats=sn)  # + shortcut = relu(bn(y0))
        sv.append((frames, y0, c0, y1, c1))
        for k in range(1, len(self.blocks)):
            blk = self.blocks[k]
            cf = blk['first_bn'].coef(sn, rows(o), train)
            sc = blk['shortcut'].conv_tc(o)  # projection of the RAW block input
            s2 = st(blk['conv1'].cout)
            y = blk['conv1'].conv_tc(o, in_bn=cf, stats=s2)
            c2 = blk['second_bn'].coef(s2, rows(y), train)
            sn = st(blk['conv2'].cout) if k + 1 < len(self.blocks) else None
            o_new = blk['conv2'].conv_tc(y, in_bn=c2, residual=sc, stats=sn)
            sv.append((o, cf, y, c2))
            o = o_new
        N = frames.shape[0]
        fl = self.flatten  # VALID convolution over the whole remaining image = a dense product on the flattened pixels
        cols = ops.round_tf32(o.reshape(N, -1))
        y = ops.empty(N, self.out_dim)
        ops.gemm(cols, self.ctx.w(fl.kernel).view(-1, self.out_dim), y, bias=self.ctx.p(fl.bias))
        self._feat = ops.relu_fwd(y, out=y)
        self._saved = (sv, cols, tuple(o.shape)) if train else None
        return self._feat

    def _backward_fused(self, dfeat):
        ctx = self.ctx
        sv, cols, oshape = self._saved
        fl = self.flatten
        d = ops.relu_bwd(self._feat, dfeat.reshape(self._feat.shape).contiguous())
        d = ops.round_tf32(d)
        ops.gemm(cols, d, ctx.g(fl.kernel).view(-1, self.out_dim), ta=True, beta=1.0)
        ops.colsum(d, ctx.g(fl.bias))
        d_o = ops.empty(*oshape)
        ops.gemm(d, ctx.w(fl.kernel).view(-1, self.out_dim), d_o.view(oshape[0], -1), tb=True)
        for k in range(len(self.blocks) - 1, 0, -1):
            blk = self.blocks[k]
            o_in, cf, y, c2 = sv[k]
            hw_y, hw_in = (y.shape[1], y.shape[2]), (o_in.shape[1], o_in.shape[2])
            blk['conv2'].wgrad_tc(y, d_o, in_bn=c2)
            s2 = ops.zeros(2 * y.shape[-1])
            dm = blk['conv2'].dgrad_tc(d_o, hw_y, mask_u=y, mask_bn=c2, stats=s2)
            du = blk['second_bn'].backward_fused(dm, y, c2, s2)
            blk['conv1'].wgrad_tc(o_in, du, in_bn=cf)
            s1 = ops.zeros(2 * o_in.shape[-1])
            d1 = blk['conv1'].dgrad_tc(du, hw_in, mask_u=o_in, mask_bn=cf, stats=s1)
            blk['shortcut'].wgrad_tc(o_in, d_o)
            d_sc = blk['shortcut'].dgrad_tc(d_o, hw_in)
            d_o = blk['first_bn'].backward_fused(d1, o_in, cf, s1, residual=d_sc)
        blk = self.blocks[0]
        frames, y0, c0, y1, c1 = sv[0]
        hw = (y0.shape[1], y0.shape[2])
        blk['conv2'].wgrad_tc(y1, d_o, in_bn=c1)
        s1 = ops.zeros(2 * y1.shape[-1])
        dm = blk['conv2'].dgrad_tc(d_o, hw, mask_u=y1, mask_bn=c1, stats=s1)
        du1 = blk['second_bn'].backward_fused(dm, y1, c1, s1)
        blk['conv1'].wgrad_tc(y0, du1, in_bn=c0)
        s0 = ops.zeros(2 * y0.shape[-1])
        dm0 = blk['conv1'].dgrad_tc(du1, hw, residual=d_o, mask_u=y0, mask_bn=c0, stats=s0)  # + the shortcut's gradient
        du0 = self.layer0_bn.backward_fused(dm0, y0, c0, s0)
        self.layer0.wgrad_tc(frames, du0)
        self._feat = self._saved = None

    def forward(self, frames, train):
        """frames [N,H,W,C] device tensor -> features [N, cnn_dense_units]."""
        self._fused = self._fused_ok()
        if self._fused:
            return self._forward_fused(frames, train)
        flow = self.layer0_bn.forward(self.layer0.forward(frames), train)
        for blk in self.blocks:
            shortcut = flow
            if 'first_bn' in blk:
                flow = blk['first_bn'].forward(flow, train)
                shortcut = blk['shortcut'].forward(shortcut)
            flow = blk['conv1'].forward(flow)
            flow = blk['second_bn'].forward(flow, train)
            flow = blk['conv2'].forward(flow)
            ops.axpy(1.0, shortcut, flow)
        y = self.flatten.forward(flow)
        self._feat = ops.relu_fwd(y, out=y)
        return self._feat.view(frames.shape[0], self.out_dim)

    def backward(self, dfeat):
        """Accumulates the gradients of every CNN variable (nothing upstream of the frames is trained)."""
        if self._fused:
            return self._backward_fused(dfeat)
        d = ops.relu_bwd(self._feat, dfeat.reshape(self._feat.shape).contiguous())
        d = self.flatten.backward(d)
        for k in range(len(self.blocks) - 1, -1, -1):
            blk = self.blocks[k]
            d_short = d  # flow + shortcut
            d = blk['conv2'].backward(d)
            d = blk['second_bn'].backward(d)
            d = blk['conv1'].backward(d)
            if 'first_bn' in blk:
                d = blk['first_bn'].backward(d)
                d_short = blk['shortcut'].backward(d_short)
            ops.axpy(1.0, d_short, d)
        d = self.layer0_bn.backward(d)
        self.layer0.backward(d, need_dx=False)
        self._feat = None

    def add_l2(self, loss_sumsq):
        """kernel_regularizer l2(1e-3) of every convolution (video.py:27, seq2seq.py:180-184): gradient += 1e-3 * w;
        loss_sumsq[0] += sum w^2 (the caller adds 1e-3 / 2 of it to the loss)."""
        ctx = self.ctx
        for name in self.kernels:
            ops.axpy(L2_SCALE, ctx.p(name).reshape(-1), ctx.g(name).reshape(-1))
            ops.sumsq(ctx.p(name).reshape(-1), loss_sumsq)


def cnn_layers(ctx, height, width, channels, cnn_type, cnn_filters, cnn_dense_units=128):
    """Factory with the dispatch of video.py:224-248."""
    if cnn_type == 'resnet_cnn':
        return ResNetCNN(ctx, height, width, channels, cnn_filters, cnn_dense_units)
    if cnn_type in ('2dconv_cnn', '3dconv_cnn'):
        raise NotImplementedError('%s (video.py) is not selected by any reference script and is not built' % cnn_type)
    raise Exception('undefined CNN, did you mean `resnet_cnn` ?')
