"""Independent pin of the oracle's recurrent core: torch.nn.LSTM (cuDNN-style LSTM equations, a third implementation
that shares no code with the oracle or the product) reproduces oracle.lstm_seq_fwd / lstm_seq_bwd once the TF layout is
mapped onto it:

  * TF LSTMCell kernel [I+H, 4H] with gate blocks (i, j, f, o)  ->  torch weight_ih / weight_hh rows (i, f, g = j, o),
  * forget_bias = 1.0 (cells.py:14-18 keeps TF's default)          ->  +1 on torch's forget-gate bias,
  * dynamic_rnn(sequence_length): zero outputs past the length, state carried to the end
                                                                  ->  pack_padded_sequence / h_n, c_n of the packed run,
  * bidirectional_dynamic_rnn: the backward direction runs on reverse_sequence(x, len) (A.2)
                                                                  ->  torch's bidirectional packed LSTM.

cell_clip = 1.0 has no torch counterpart: the cases keep |c| < 1 (asserted), so the clip is the identity; the clip
itself is covered by the finite-difference checks.  Gradients: torch autograd against the oracle's analytic backward,
every entry."""
import numpy as np
import pytest
import torch

from oracle import avsr_oracle as O


def tf_to_torch(W, b, I, H):
    blocks = {'i': slice(0, H), 'j': slice(H, 2 * H), 'f': slice(2 * H, 3 * H), 'o': slice(3 * H, 4 * H)}
    order = ['i', 'f', 'j', 'o']  # torch: input, forget, cell (g), output
    w_ih = np.concatenate([W[:I, blocks[g]].T for g in order], axis=0)
    w_hh = np.concatenate([W[I:, blocks[g]].T for g in order], axis=0)
    bias = np.concatenate([b[blocks[g]] + (O.FORGET_BIAS if g == 'f' else 0.0) for g in order])
    return w_ih, w_hh, bias


def torch_grad_to_tf(g_ih, g_hh, g_b, I, H):
    """inverse mapping for gradients: rows (i, f, g, o) of the torch matrices -> columns (i, j, f, o) of the TF kernel."""
    idx = {'i': slice(0, H), 'f': slice(H, 2 * H), 'j': slice(2 * H, 3 * H), 'o': slice(3 * H, 4 * H)}
    dW = np.zeros((I + H, 4 * H))
    db = np.zeros(4 * H)
    for k, g in enumerate(['i', 'j', 'f', 'o']):
        dW[:I, k * H:(k + 1) * H] = g_ih[idx[g]].T
        dW[I:, k * H:(k + 1) * H] = g_hh[idx[g]].T
        db[k * H:(k + 1) * H] = g_b[idx[g]]
    return dW, db


def make_lstm(W, b, I, H, bidirectional=False, Wb=None, bb=None):
    lstm = torch.nn.LSTM(I, H, batch_first=True, bidirectional=bidirectional).double()
    with torch.no_grad():
        for sfx, (Wx, bx) in (('', (W, b)),) + ((('_reverse', (Wb, bb)),) if bidirectional else ()):
            w_ih, w_hh, bias = tf_to_torch(Wx, bx, I, H)
            getattr(lstm, 'weight_ih_l0' + sfx).copy_(torch.tensor(w_ih))
            getattr(lstm, 'weight_hh_l0' + sfx).copy_(torch.tensor(w_hh))
            getattr(lstm, 'bias_ih_l0' + sfx).copy_(torch.tensor(bias))
            getattr(lstm, 'bias_hh_l0' + sfx).zero_()
    return lstm


@pytest.mark.parametrize('B,T,I,H,seed,scale', [(4, 9, 5, 6, 0, 0.3), (3, 1, 2, 3, 1, 0.3), (6, 17, 8, 16, 2, 0.12)])
def test_lstm_layer_matches_torch_lstm(B, T, I, H, seed, scale):
    rng = np.random.default_rng(seed)
    W = scale * rng.standard_normal((I + H, 4 * H))
    b = 0.1 * rng.standard_normal(4 * H)
    x = 0.5 * rng.standard_normal((B, T, I))
    lens = rng.integers(1, T + 1, B)
    lens[0] = T
    x *= (np.arange(T)[None, :, None] < lens[:, None, None])
    outs, (c_fin, h_fin), cache = O.lstm_seq_fwd(x, lens, W, b)
    assert all(np.abs(cc[6]).max() < O.CELL_CLIP for cc, *_ in cache[0])  # c_raw never reaches the clip
    dout, dc, dh = (rng.standard_normal(a.shape) for a in (outs, c_fin, h_fin))
    dx, dW, db, _ = O.lstm_seq_bwd(dout, (dc, dh), cache)

    lstm = make_lstm(W, b, I, H)
    xt = torch.tensor(x, requires_grad=True)
    packed = torch.nn.utils.rnn.pack_padded_sequence(xt, torch.tensor(lens), batch_first=True, enforce_sorted=False)
    out_p, (h_n, c_n) = lstm(packed)
    out_t, _ = torch.nn.utils.rnn.pad_packed_sequence(out_p, batch_first=True, total_length=T)
    assert np.allclose(outs, out_t.detach().numpy(), atol=1e-12)   # zero past the length on both sides
    assert np.allclose(h_fin, h_n[0].detach().numpy(), atol=1e-12)  # state at each row's last valid step
    assert np.allclose(c_fin, c_n[0].detach().numpy(), atol=1e-12)
    ((out_t * torch.tensor(dout)).sum() + (h_n[0] * torch.tensor(dh)).sum() + (c_n[0] * torch.tensor(dc)).sum()).backward()
    dW_t, db_t = torch_grad_to_tf(lstm.weight_ih_l0.grad.numpy(), lstm.weight_hh_l0.grad.numpy(),
                                  lstm.bias_ih_l0.grad.numpy(), I, H)
    assert np.allclose(dx, xt.grad.numpy(), atol=1e-11)
    assert np.allclose(dW, dW_t, atol=1e-10) and np.allclose(db, db_t, atol=1e-10)


def test_bidirectional_layer_matches_torch():
    """One layer per direction (encoder.py:90-143 with L = 1 stacks): reverse_sequence semantics of the backward run."""
    rng = np.random.default_rng(7)
    B, T, I, H = 5, 11, 4, 6
    Wf, Wb = (0.3 * rng.standard_normal((I + H, 4 * H)) for _ in range(2))
    bf, bb = (0.1 * rng.standard_normal(4 * H) for _ in range(2))
    x = 0.5 * rng.standard_normal((B, T, I))
    lens = np.array([11, 3, 7, 1, 10])
    x *= (np.arange(T)[None, :, None] < lens[:, None, None])
    out, (sf, sb), cache = O.birnn_fwd(x, lens, [(Wf, bf)], [(Wb, bb)])
    lstm = make_lstm(Wf, bf, I, H, bidirectional=True, Wb=Wb, bb=bb)
    xt = torch.tensor(x, requires_grad=True)
    packed = torch.nn.utils.rnn.pack_padded_sequence(xt, torch.tensor(lens), batch_first=True, enforce_sorted=False)
    out_p, (h_n, c_n) = lstm(packed)
    out_t, _ = torch.nn.utils.rnn.pad_packed_sequence(out_p, batch_first=True, total_length=T)
    assert np.allclose(out, out_t.detach().numpy(), atol=1e-12)
    assert np.allclose(sf[0][1], h_n[0].detach().numpy(), atol=1e-12) and np.allclose(sb[0][1], h_n[1].detach().numpy(), atol=1e-12)
    assert np.allclose(sf[0][0], c_n[0].detach().numpy(), atol=1e-12) and np.allclose(sb[0][0], c_n[1].detach().numpy(), atol=1e-12)
    dout = rng.standard_normal(out.shape)
    dx, gf, gb = O.birnn_bwd(dout, None, None, cache)
    (out_t * torch.tensor(dout)).sum().backward()
    assert np.allclose(dx, xt.grad.numpy(), atol=1e-11)
    dWf, dbf = torch_grad_to_tf(lstm.weight_ih_l0.grad.numpy(), lstm.weight_hh_l0.grad.numpy(), lstm.bias_ih_l0.grad.numpy(), I, H)
    dWb, dbb = torch_grad_to_tf(lstm.weight_ih_l0_reverse.grad.numpy(), lstm.weight_hh_l0_reverse.grad.numpy(),
                                lstm.bias_ih_l0_reverse.grad.numpy(), I, H)
    assert np.allclose(gf[0][0], dWf, atol=1e-10) and np.allclose(gf[0][1], dbf, atol=1e-10)
    assert np.allclose(gb[0][0], dWb, atol=1e-10) and np.allclose(gb[0][1], dbb, atol=1e-10)


@pytest.mark.parametrize('kind', ['luong', 'scaled_luong'])
def test_luong_attention_matches_torch_sdpa(kind):
    """score -> masked softmax -> context of the Luong family (attention.py:55-72) is scaled dot-product
    attention with scale = attention_g (1 for plain luong) and a key mask from memory_sequence_length:
    torch.nn.functional.scaled_dot_product_attention is an independent implementation of exactly that.  The queries
    are the cell outputs of the oracle's own AttentionWrapper run."""
    import torch.nn.functional as F
    rng = np.random.default_rng(3)
    B, T, Dx, H, Tm, Dm = 4, 6, 5, 8, 9, 7
    mem_len = np.array([9, 4, 1, 6])
    memory = rng.standard_normal((B, Tm, Dm))
    spec = O.AttnSpec(kind=kind, memory=memory, mem_len=mem_len, Wm=0.4 * rng.standard_normal((Dm, H)),
                      Wl=0.3 * rng.standard_normal((H + Dm, H)), g=np.asarray(1.7) if kind == 'scaled_luong' else None)
    W = 0.3 * rng.standard_normal((Dx + H + H, 4 * H))
    b = 0.1 * rng.standard_normal(4 * H)
    x = rng.standard_normal((B, T, Dx))
    lens = np.array([6, 6, 3, 5])
    r = O.attn_rnn_fwd(x, lens, W, b, [spec])
    values, keys, mask = O._prepare_memory(spec)
    q = torch.tensor(r['cell_outputs'])                       # [B,T,H] (zero past the length)
    key_mask = torch.tensor(mask > 0)[:, None, :].expand(B, T, Tm)
    ctx = F.scaled_dot_product_attention(q, torch.tensor(keys), torch.tensor(values), attn_mask=key_mask,
                                         scale=float(spec.g) if spec.g is not None else 1.0)
    step_mask = (np.arange(T)[None, :] < lens[:, None])[:, :, None]
    assert np.allclose(r['contexts'][0], ctx.numpy() * step_mask, atol=1e-12)
    # alignments: rows are distributions over the valid memory positions only
    a = r['alignments'][0]
    assert np.allclose(a.sum(-1), step_mask[:, :, 0].astype(float), atol=1e-12)
    assert not (a * (1 - mask)[:, None, :]).any()


def test_sequence_loss_matches_torch_cross_entropy():
    """seq2seq.sequence_loss(average_across_timesteps and batch) with sequence_mask weights (seq2seq.py:142-171, A.5)
    against torch's cross_entropy + autograd."""
    import torch.nn.functional as F
    rng = np.random.default_rng(11)
    B, T, V = 5, 7, 31
    logits = rng.standard_normal((B, T, V)) * 2
    tgt = rng.integers(0, V, (B, T))
    lens = np.array([7, 2, 5, 1, 6])
    loss, dlogits = O.sequence_loss_fwd_bwd(logits, tgt, lens)
    zt = torch.tensor(logits, requires_grad=True)
    w = torch.tensor((np.arange(T)[None, :] < lens[:, None]).astype(np.float64))
    ce = F.cross_entropy(zt.reshape(-1, V), torch.tensor(tgt).reshape(-1), reduction='none').reshape(B, T)
    lt = (ce * w).sum() / (w.sum() + 1e-12)
    lt.backward()
    assert abs(loss - float(lt.detach())) < 1e-12
    assert np.allclose(dlogits, zt.grad.numpy(), atol=1e-14)


@pytest.mark.parametrize('optimiser', ['Adam', 'AdamW', 'Momentum'])
def test_optimisers_match_torch_optim(optimiser):
    """clip_by_global_norm + the optimisers of seq2seq.py:195-219 against torch.optim over five steps.
    TF-Adam folds the bias corrections into the step size and adds epsilon to sqrt(v) un-corrected; torch corrects v
    first: the two differ only through epsilon (1e-8 against gradients of order 1), hence the 1e-6 tolerance on the
    updates.  tf.contrib's AdamW decays by weight_decay un-scaled by the learning rate = torch's AdamW with
    weight_decay / lr.  torch's clip_grad_norm_ divides by (norm + 1e-6)."""
    rng = np.random.default_rng(5)
    lr, wd, clip = 1e-2, 1e-3, 1.0
    P = {'a': rng.standard_normal((4, 3)), 'b': rng.standard_normal(5)}
    P0 = {k: v.copy() for k, v in P.items()}
    grads = [{k: rng.standard_normal(v.shape) * (3.0 if s % 2 else 0.2) for k, v in P.items()} for s in range(5)]
    m = {k: np.zeros_like(v) for k, v in P.items()}
    v_ = {k: np.zeros_like(v) for k, v in P.items()}
    tp = {k: torch.nn.Parameter(torch.tensor(v)) for k, v in P.items()}
    opt = {'Adam': lambda: torch.optim.Adam(tp.values(), lr=lr, eps=1e-8),
           'AdamW': lambda: torch.optim.AdamW(tp.values(), lr=lr, eps=1e-8, weight_decay=wd / lr),
           'Momentum': lambda: torch.optim.SGD(tp.values(), lr=lr, momentum=0.9)}[optimiser]()
    for s, G in enumerate(grads):
        gn = O.clip_and_adam(P, G, m, v_, s, lr, clip=clip, warmup_steps=None, optimiser=optimiser, weight_decay=wd)
        for k in tp:
            tp[k].grad = torch.tensor(G[k])
        gn_t = torch.nn.utils.clip_grad_norm_(tp.values(), clip)
        assert abs(gn - float(gn_t)) < 1e-12  # both return the pre-clip norm
        opt.step()
    for k in P:
        moved = np.abs(P[k] - P0[k]).max()
        assert moved > 1e-3
        assert np.abs(P[k] - tp[k].detach().numpy()).max() <= 2e-6 * max(moved, 1.0), k
