"""Independent pin of the oracle's visual front-end: the same network expressed with PyTorch's own primitives on the CPU
(F.conv2d with TensorFlow's asymmetric SAME padding applied by hand, F.batch_norm in training mode, autograd for every
gradient) must agree with oracle.resnet_cnn_fwd / resnet_cnn_bwd, which are written as im2col products with analytic
backward.  (TF itself cannot be installed here; torch's convolution and batch-norm are a third implementation of the same
published semantics.)"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import avsr_oracle as O


def tf_same_pad(x, k, stride):
    """x [N,C,H,W] -> padded like tf.layers.conv2d(padding='SAME'): the odd pixel goes to the bottom / right."""
    H, W = x.shape[2:]
    _, pt, pb = O.same_padding(H, k, stride)
    _, pl, pr = O.same_padding(W, k, stride)
    return F.pad(x, (pl, pr, pt, pb))


def torch_resnet(P, frames, filters):
    def conv(x, name, k=3, stride=1, padding='SAME'):
        w = P['CNN/' + name + '/kernel'].permute(3, 2, 0, 1)  # [kh,kw,Ci,Co] -> [Co,Ci,kh,kw]
        if padding == 'SAME':
            x = tf_same_pad(x, k, stride)
        return F.conv2d(x, w, P['CNN/' + name + '/bias'], stride=stride)

    def bnr(x, name):
        y = F.batch_norm(x, None, None, P['CNN/' + name + '/gamma'], P['CNN/' + name + '/beta'], training=True,
                         eps=O.CNN_BN_EPS)
        return torch.relu(y)

    x = frames.permute(0, 3, 1, 2)  # NHWC -> NCHW
    flow = bnr(conv(x, 'layer0'), 'layer0_bn')
    for k, _ in enumerate(filters):
        n = 'res_block_%d' % k
        shortcut = flow
        if k > 0:
            flow = bnr(flow, n + '_first_bn')
            shortcut = conv(shortcut, n + '_shortcut', k=1, stride=2)
        flow = conv(flow, n + '_conv1', stride=1 if k == 0 else 2)
        flow = bnr(flow, n + '_second_bn')
        flow = conv(flow, n + '_conv2') + shortcut
    ksz = flow.shape[2]
    return torch.relu(conv(flow, 'flatten', k=ksz, padding='VALID')).flatten(1)


def make_params(rng, hw, ch, filters, dense):
    P = {}

    def conv(name, k, ci, co):
        P['CNN/' + name + '/kernel'] = rng.standard_normal((k, k, ci, co)) * np.sqrt(2.0 / (k * k * ci))
        P['CNN/' + name + '/bias'] = 0.1 * rng.standard_normal(co)

    def bn(name, c):
        P['CNN/' + name + '/gamma'] = 1 + 0.2 * rng.standard_normal(c)
        P['CNN/' + name + '/beta'] = 0.1 * rng.standard_normal(c)

    conv('layer0', 3, ch, filters[0])
    bn('layer0_bn', filters[0])
    cin, size = filters[0], hw
    for k, f in enumerate(filters):
        n = 'res_block_%d' % k
        if k > 0:
            bn(n + '_first_bn', cin)
            conv(n + '_shortcut', 1, cin, f)
            size = -(-size // 2)
        conv(n + '_conv1', 3, cin, f)
        bn(n + '_second_bn', f)
        conv(n + '_conv2', 3, f, f)
        cin = f
    conv('flatten', size, cin, dense)
    return P


@pytest.mark.parametrize('hw,filters,dense', [(36, (8, 16, 32, 64), 128), (13, (3, 5, 4, 6), 7), (8, (2, 3, 4, 5), 6)])
def test_oracle_cnn_matches_torch_autograd(hw, filters, dense):
    """36 -> 18 -> 9 -> 5 exercises the asymmetric SAME padding of the stride-2 layers (even sizes pad 0 / 1, odd 1 / 1)."""
    rng = np.random.default_rng(hw)
    P = make_params(rng, hw, 3, filters, dense)
    frames = rng.uniform(-1, 1, (4, hw, hw, 3))
    w = rng.standard_normal((4, dense))
    feat, cache, stats = O.resnet_cnn_fwd(P, frames, filters)
    dframes, G = O.resnet_cnn_bwd(w, cache)

    Pt = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in P.items()}
    xt = torch.tensor(frames, dtype=torch.float64, requires_grad=True)
    ft = torch_resnet(Pt, xt, filters)
    (ft * torch.tensor(w)).sum().backward()
    assert feat.shape == tuple(ft.shape)
    assert np.allclose(feat, ft.detach().numpy(), rtol=1e-9, atol=1e-11)
    assert np.allclose(dframes, xt.grad.numpy(), rtol=1e-7, atol=1e-10)
    assert set(G) == set(P)
    for name, g in G.items():
        ref = Pt[name].grad.numpy()
        assert np.allclose(g, ref, rtol=1e-7, atol=1e-9 * max(1.0, np.abs(ref).max())), name
    # batch statistics handed to the moving averages: biased variance over N, H, W
    a0, _ = O.conv2d_fwd(frames, P['CNN/layer0/kernel'], P['CNN/layer0/bias'])
    mean, var = stats['CNN/layer0_bn']
    assert np.allclose(mean, a0.mean(axis=(0, 1, 2))) and np.allclose(var, a0.var(axis=(0, 1, 2)))


@pytest.mark.parametrize('H,W,k,stride', [(36, 36, 3, 2), (18, 18, 3, 2), (9, 9, 3, 2), (9, 7, 3, 1), (18, 18, 1, 2)])
def test_oracle_conv_same_padding_matches_torch(H, W, k, stride):
    rng = np.random.default_rng(H * W + k + stride)
    x = rng.standard_normal((2, H, W, 3))
    K = rng.standard_normal((k, k, 3, 4))
    b = rng.standard_normal(4)
    y, _ = O.conv2d_fwd(x, K, b, stride, 'SAME')
    xt = tf_same_pad(torch.tensor(x).permute(0, 3, 1, 2), k, stride)
    yt = F.conv2d(xt, torch.tensor(K).permute(3, 2, 0, 1), torch.tensor(b), stride=stride).permute(0, 2, 3, 1)
    assert y.shape == tuple(yt.shape) == (2, -(-H // stride), -(-W // stride), 4)
    assert np.allclose(y, yt.numpy(), rtol=1e-10, atol=1e-12)
