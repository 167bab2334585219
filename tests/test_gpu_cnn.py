"""GPU parity of the visual front-end (avsr/video.py resnet_cnn) against the oracle: the im2col /
col2im pair, the CNN alone (features and every gradient), and the whole model with lip crops as input."""
import numpy as np
import pytest
import torch

from oracle import avsr_oracle as O
from tests.helpers import (add_aus, cast_batch, config_hparams, oracle_hparams, synthetic_batch, to_data_sequences,
                           to_image_sequences)
from tests.test_gpu_model import close, tensor_cores  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('H,W,C,k,stride,padding', [(9, 9, 3, 3, 1, 'SAME'), (9, 7, 4, 3, 2, 'SAME'),
                                                     (36, 36, 3, 3, 2, 'SAME'), (18, 18, 8, 1, 2, 'SAME'),
                                                     (5, 5, 6, 5, 1, 'VALID'), (8, 8, 2, 3, 2, 'VALID')])
def test_im2col_col2im(H, W, C, k, stride, padding):
    from avsr_tf1_b200 import ops
    old = ops.set_tensor_cores(False)
    try:
        rng = np.random.default_rng(H + W + C + k + stride)
        x = rng.standard_normal((3, H, W, C)).astype(np.float32)
        cols_ref, geom_ref = O.im2col(x, k, k, stride, padding)
        cols, geom = ops.im2col(torch.from_numpy(x).cuda(), k, k, stride, padding)
        assert np.array_equal(cols.cpu().numpy(), cols_ref)
        d = rng.standard_normal(cols_ref.shape).astype(np.float32)
        dx_ref = O.col2im(d.astype(np.float64), geom_ref, k, k, stride)
        dx = ops.col2im(torch.from_numpy(d).cuda(), geom)
        close(dx, dx_ref, 1e-6, 'col2im')
    finally:
        ops.set_tensor_cores(old)


def test_resnet_cnn_alone(tensor_cores):
    """36x36x3 crops, the reference's filters (8, 16, 32, 64) -> 128 features; training-mode batch statistics."""
    from avsr_tf1_b200.layers import BuildContext
    from avsr_tf1_b200.params import ParamStore
    from avsr_tf1_b200.video import ResNetCNN
    ctx = BuildContext()
    cnn = ResNetCNN(ctx, 36, 36, 3)
    ctx.store = ParamStore(ctx.specs, device='cuda', with_optimizer=True)
    ctx.store.initialize(7)
    rng = np.random.default_rng(0)
    for s in ctx.specs:  # non-trivial BN affine parameters and biases
        if s.name.endswith(('gamma', 'beta', 'bias')):
            ctx.store.p(s.name).add_(torch.from_numpy(0.2 * rng.standard_normal(s.shape).astype(np.float32)).cuda())
    ctx.store.sync_tf32()
    P = {k: v.astype(np.float64) for k, v in ctx.store.to_numpy('p').items()}
    frames = rng.uniform(-1, 1, (6, 36, 36, 3)).astype(np.float32)
    w = rng.standard_normal((6, 128)).astype(np.float32)
    feat_ref, cache, stats = O.resnet_cnn_fwd(P, frames.astype(np.float64))
    _, G_ref = O.resnet_cnn_bwd(w.astype(np.float64), cache)
    ctx.store.grad.zero_()
    feat = cnn.forward(torch.from_numpy(frames).cuda(), train=True)
    rt = 5e-3 if tensor_cores else 1e-4  # 13 tf32 products deep
    close(feat, feat_ref, rt, 'cnn features')
    cnn.backward(torch.from_numpy(w).cuda())
    G = ctx.store.to_numpy('g')
    gmax = max(np.abs(g).max() for g in G_ref.values())
    for name, g_ref in G_ref.items():
        scale = max(np.abs(g_ref).max(), 1e-3 * gmax)
        got = G[name].astype(np.float64)
        err = np.abs(got - g_ref).max() / scale
        if not tensor_cores:
            assert err <= 1e-3, f'{name}: gradient scaled error {err:.3e}'  # exact mode pins the algorithm
        elif np.abs(g_ref).max() > 1e-3 * gmax:
            # tf32-rounded operands flip a few ReLUs of the deep layers, and with 6 frames (150 positions in the last
            # block) one flip moves single entries visibly: direction and size of each gradient tensor instead
            cos = float((got * g_ref).sum() / (np.linalg.norm(got) * np.linalg.norm(g_ref) + 1e-30))
            ratio = float(np.linalg.norm(got) / (np.linalg.norm(g_ref) + 1e-30))
            assert cos >= 0.97 and abs(ratio - 1.0) <= 0.1, f'{name}: cosine {cos:.4f}, norm ratio {ratio:.3f}'
    # moving statistics: momentum 0.98 (video.py:10)
    mean, var = stats['CNN/layer0_bn']
    close(ctx.store.p('CNN/layer0_bn/moving_mean'), 0.02 * mean, 2e-3, 'moving mean')
    close(ctx.store.p('CNN/layer0_bn/moving_variance'), 0.98 + 0.02 * var, 1e-4, 'moving variance')
    # inference mode uses them
    feat_eval = cnn.forward(torch.from_numpy(frames).cuda(), train=False)
    P2 = {k: v.astype(np.float64) for k, v in ctx.store.to_numpy('p').items()}
    feat_eval_ref, _, _ = O.resnet_cnn_fwd(P2, frames.astype(np.float64), train=False)
    close(feat_eval, feat_eval_ref, rt, 'cnn features (inference)')


@pytest.mark.parametrize('cfg,over', [(3, {}), (5, {}), (4, dict(regress_aus=True)),
                                      (5, dict(use_dropout=True, sampling_probability_outputs=0.0))])
def test_model_with_cnn_front_end(cfg, over, tensor_cores):
    """video_processing='resnet_cnn' (run_video.py:34, run_audiovisual.py): lip crops -> CNN -> video encoder -> ...;
    loss (incl. the conv L2 term), encoder states and every gradient incl. the CNN's."""
    from avsr_tf1_b200.seq2seq import Seq2SeqModel
    hp = config_hparams(cfg, video_processing='resnet_cnn', cnn_filters=(4, 8, 8, 16), cnn_dense_units=32, **over)
    batch = to_image_sequences(synthetic_batch(hp, B=3, Ta=24, Tv=6, L=5, ragged=True), hw=12)
    if hp.regress_aus:
        add_aus(batch)
    ds = to_data_sequences(batch)
    model = Seq2SeqModel(ds, 'train', hp, seed=2001)
    P = {k: v.astype(np.float64) for k, v in model.store.to_numpy('p').items()}
    om = O.OracleModel(oracle_hparams(hp, model), P)
    loss_ref, G_ref, rec = om.loss_and_grads(cast_batch(batch, np.float64))
    model.feed(ds)
    model._set_step_scalars()
    model.forward_backward()
    model.finish_gradients()
    loss, gnorm = model.fetch_scalars()
    assert abs(loss - loss_ref) <= 1e-3 * abs(loss_ref), (loss, loss_ref)
    out_ref, (c_ref, h_ref) = rec['enc']['video']
    d = model._video_encoder.get_data()
    rt = 3e-3 if tensor_cores else 1e-3
    close(d.outputs.transpose(0, 1), out_ref, rt, 'video encoder outputs')
    close(d.final_state[1], h_ref, rt, 'video final h')
    G = model.store.to_numpy('g')
    gn_ref = O.global_norm(G_ref)
    assert abs(gnorm - gn_ref) <= 1e-2 * gn_ref, (gnorm, gn_ref)
    gmax = max(np.abs(g).max() for g in G_ref.values())
    for name, g_ref in G_ref.items():
        scale = max(np.abs(g_ref).max(), 1e-3 * gmax)
        err = np.abs(G[name].astype(np.float64) - g_ref).max() / scale
        # exact-fp32 mode pins the algorithm at 1e-3 everywhere; in tensor-core mode the operands of the wide
        # convolutions are tf32-rounded and the BN / bias gradients are cancellation-heavy sums over all pixels of a tiny
        # batch, where a few flipped ReLUs show in single entries
        if tensor_cores and name.startswith('CNN/'):
            continue  # judged together below
        assert err <= (3e-2 if tensor_cores else 1e-3), f'{name}: gradient scaled error {err:.3e}'
    if tensor_cores:
        # the CNN's gradients in tensor-core mode: 3 utterances x 6 crops of 12 x 12 pixels leave a handful of positions
        # per channel in the deep layers, so one ReLU flipped by a tf32-rounded operand moves a 4-entry gamma gradient by
        # 20 %; direction and size of the whole CNN gradient are what the tiny batch can pin (exact mode pins every entry)
        names = [n for n in G_ref if n.startswith('CNN/')]
        got = np.concatenate([G[n].astype(np.float64).reshape(-1) for n in names])
        ref = np.concatenate([G_ref[n].reshape(-1) for n in names])
        cos = float((got * ref).sum() / (np.linalg.norm(got) * np.linalg.norm(ref) + 1e-30))
        ratio = float(np.linalg.norm(got) / (np.linalg.norm(ref) + 1e-30))
        assert np.isfinite(got).all() and cos >= 0.95 and abs(ratio - 1.0) <= 0.15, (cos, ratio)
    # three optimiser steps run, and inference works on crops
    for _ in range(2):
        l2, _ = model.train_step(ds)
        assert np.isfinite(l2)
    hp_eval = config_hparams(cfg, video_processing='resnet_cnn', cnn_filters=(4, 8, 8, 16), cnn_dense_units=32,
                             decoding_algorithm='greedy', **over)
    ev = Seq2SeqModel(ds, 'evaluate', hp_eval, share_params_with=model)
    ids = ev.predict(ds)
    assert ids.shape[0] == 3


@pytest.mark.parametrize('N,H,Ci,Co,k,stride', [
    (5, 36, 3, 8, 3, 1), (3, 36, 8, 8, 3, 1), (4, 36, 8, 16, 1, 2), (4, 36, 8, 16, 3, 2), (7, 18, 16, 16, 3, 1),
    (5, 18, 16, 32, 1, 2), (5, 18, 16, 32, 3, 2), (9, 9, 32, 32, 3, 1), (9, 9, 32, 64, 1, 2), (9, 9, 32, 64, 3, 2),
    (37, 5, 64, 64, 3, 1), (2, 12, 4, 8, 3, 2), (300, 9, 32, 32, 3, 1),
])
def test_tensor_core_convolutions_against_torch(N, H, Ci, Co, k, stride):
    """csrc/conv_mma.cu against torch.nn.functional.conv2d (fp32 reference of the same op, TF's SAME padding applied by hand)
    and its autograd: forward with bias + residual + fused BN statistics, the input gradient (stride 1: flipped kernel;
    stride 2: zero-stuffed dy) and the weight gradient, on every layer shape of the reference's front-end.  Operands are
    tf32-rounded by the kernels: 2e-3 of the tensor's scale."""
    import torch.nn.functional as F
    from avsr_tf1_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(N * 1000 + H * 10 + Ci + Co + k + stride)
    x = torch.randn(N, H, H, Ci, device='cuda', generator=g)
    w = torch.randn(k, k, Ci, Co, device='cuda', generator=g) / (k * k * Ci) ** 0.5
    b = torch.randn(Co, device='cuda', generator=g)
    Ho, Wo, pt, pl = ops.conv_geometry(H, H, k, k, stride, 'SAME')
    _, _, pb = ops.same_padding(H, k, stride)
    res = torch.randn(N, Ho, Wo, Co, device='cuda', generator=g)
    old_tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        xr = x.permute(0, 3, 1, 2).clone().requires_grad_(True)
        wr = w.permute(3, 2, 0, 1).clone().requires_grad_(True)
        yr = F.conv2d(F.pad(xr, (pl, pb, pt, pb)), wr, b, stride=stride)
        dy = torch.randn(N, Ho, Wo, Co, device='cuda', generator=g)
        yr.backward(dy.permute(0, 3, 1, 2))
    finally:
        torch.backends.cudnn.allow_tf32 = old_tf32
    y_ref = yr.detach().permute(0, 2, 3, 1) + res
    stats = torch.zeros(2 * Co, device='cuda')
    y = ops.conv2d_tc(x, w.reshape(-1, Co), b, k, k, stride, pt, pl, Ho, Wo, residual=res, stats=stats)
    close(y, y_ref.cpu().numpy(), 2e-3, 'conv forward')
    y2 = y_ref.reshape(-1, Co).double()
    close(stats[:Co], y2.sum(0).cpu().numpy(), 2e-3, 'sum of y')
    close(stats[Co:], (y2 * y2).sum(0).cpu().numpy(), 2e-3, 'sum of y^2')
    dW, db = torch.zeros(k * k * Ci, Co, device='cuda'), torch.zeros(Co, device='cuda')
    ops.conv2d_wgrad_tc(x, dy, k, k, stride, 'SAME', dW, dbias=db)
    close(dW, wr.grad.permute(2, 3, 1, 0).reshape(-1, Co).cpu().numpy(), 2e-3, 'weight gradient')
    close(db, dy.double().sum((0, 1, 2)).cpu().numpy(), 1e-4, 'bias gradient')
    if Ci % 8 == 0:
        wt = w.flip(0, 1).permute(0, 1, 3, 2).contiguous().view(-1, Ci)
        dx = ops.conv2d_tc(dy, wt, None, k, k, 1, k - 1 - pt, k - 1 - pl, H, H, in_dilation=stride)
        close(dx, xr.grad.permute(0, 2, 3, 1).cpu().numpy(), 2e-3, 'input gradient')


@pytest.mark.parametrize('N,H,C', [(6, 36, 8), (10, 18, 16), (40, 9, 32)])
def test_fused_batch_norm_relu_block_against_torch(N, H, C):
    """conv -> batch_norm_relu -> conv + shortcut (res_block_0 of video.py:57-92 with skip_bn) with the BN-ReLU fused into
    the convolutions around it (statistics in the producer's epilogue, apply in the consumer's loads, ReLU mask and backward
    statistics in the epilogue of the input-gradient convolution) against torch autograd in fp32.  The torch chain starts
    from the first convolution's output as the kernel produced it (that convolution is checked on its own above), so both
    sides normalise - and cut at zero - the same numbers; what remains is the tf32 rounding of the operands."""
    import torch.nn.functional as F
    from avsr_tf1_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(N + H + C)
    rnd = lambda *s: torch.randn(*s, device='cuda', generator=g)
    x, wa, wb = rnd(N, H, H, C), rnd(3, 3, C, C) / (9 * C) ** 0.5, rnd(3, 3, C, C) / (9 * C) ** 0.5
    ba, bb, gamma, beta, wout = 0.1 * rnd(C), 0.1 * rnd(C), 1 + 0.2 * rnd(C), 0.2 * rnd(C), rnd(N, H, H, C)
    eps = 1e-5
    count = N * H * H
    s0 = torch.zeros(2 * C, device='cuda')
    y0k = ops.conv2d_tc(x, wa.reshape(-1, C), ba, 3, 3, 1, 1, 1, H, H, stats=s0)
    old_tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        y0, wbr, bbr, gr, br = [t.clone().requires_grad_(True) for t in (y0k, wb, bb, gamma, beta)]
        conv = lambda t, w, b: F.conv2d(t.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), b, padding=1).permute(0, 2, 3, 1)
        mean, var = y0.mean((0, 1, 2)), y0.var((0, 1, 2), unbiased=False)
        z = torch.relu((y0 - mean) / torch.sqrt(var + eps) * gr + br)
        o = conv(z, wbr, bbr) + z
        (o * wout).sum().backward()
    finally:
        torch.backends.cudnn.allow_tf32 = old_tf32

    def l2(got, want, what, tol=3e-3):
        err = float(torch.linalg.norm(got.double() - want.double()) / (torch.linalg.norm(want.double()) + 1e-30))
        assert np.isfinite(err) and err <= tol, f'{what}: relative L2 error {err:.3e}'

    l2(s0[:C] / count, mean.detach(), 'fused mean', 1e-4)
    l2(s0[C:] / count - (s0[:C] / count) ** 2, var.detach(), 'fused variance', 1e-4)
    mm, mv = torch.zeros(C, device='cuda'), torch.ones(C, device='cuda')
    c0 = ops.bn_finalize(s0, count, gamma, beta, eps, 0.98, mm, mv)
    l2(mm, 0.02 * mean.detach(), 'moving mean', 1e-4)
    l2(mv, 0.98 + 0.02 * var.detach(), 'moving variance', 1e-4)
    ok = ops.conv2d_tc(y0k, wb.reshape(-1, C), bb, 3, 3, 1, 1, 1, H, H, in_bn=c0, residual=y0k, res_bn=c0)
    l2(ok, o.detach(), 'block output')
    dWb = torch.zeros(9 * C, C, device='cuda')
    ops.conv2d_wgrad_tc(y0k, wout, 3, 3, 1, 'SAME', dWb, in_bn=c0)
    l2(dWb, wbr.grad.reshape(-1, C), 'dW of the second convolution')
    s2 = torch.zeros(2 * C, device='cuda')
    flipT = lambda w: w.flip(0, 1).permute(0, 1, 3, 2).contiguous().view(-1, C)
    dm = ops.conv2d_tc(wout, flipT(wb), None, 3, 3, 1, 1, 1, H, H, residual=wout, mask_u=y0k, mask_bn=c0, stats=s2)
    l2(s2[:C], br.grad, 'dbeta')
    l2(s2[C:], gr.grad, 'dgamma')
    du0 = ops.bn_relu_bwd_apply(dm, y0k, c0, s2, count)
    l2(du0, y0.grad, 'gradient wrt the batch norm input')
    res = rnd(N, H, H, C)
    du1 = ops.bn_relu_bwd_apply(dm, y0k, c0, s2, count, residual=res)
    l2(du1, y0.grad + res, 'gradient wrt the batch norm input + residual')
    # inference coefficients from the moving statistics
    ce = ops.bn_coef_eval(gamma, beta, mm, mv, eps)
    ze = ops.conv2d_tc(y0k, torch.eye(C, device='cuda').reshape(1, 1, C, C).reshape(-1, C), None, 1, 1, 1, 0, 0, H, H, in_bn=ce)
    l2(ze, torch.relu((y0k - mm) / torch.sqrt(mv + eps) * gamma + beta), 'inference-mode BN-ReLU through a 1x1 identity')
