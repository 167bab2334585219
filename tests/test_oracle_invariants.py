"""Algebraic invariants of the oracle (validation by independent means)."""
import numpy as np

from oracle import avsr_oracle as O


def rnd(seed, *shape):
    return np.random.default_rng(seed).standard_normal(shape)


def test_lstm_cell_hand_case():
    """H = 1, hand-computed: z = 0 everywhere -> i = o = 0.5, f = sigmoid(1), j = 0."""
    W, b = np.zeros((2, 4)), np.zeros(4)
    h, c, _ = O.lstm_cell(np.array([[0.3, -0.2]]), np.array([[0.5]]), W, b)
    f = 1 / (1 + np.exp(-1.0))
    assert np.allclose(c, f * 0.5) and np.allclose(h, 0.5 * np.tanh(f * 0.5))
    # cell clip at 1.0 and gate order i, j, f, o
    W = np.zeros((2, 4)); b = np.array([50.0, 50.0, 50.0, 50.0])
    h, c, _ = O.lstm_cell(np.array([[0.0, 0.0]]), np.array([[1.0]]), W, b)
    assert np.allclose(c, 1.0) and np.allclose(h, np.tanh(1.0))


def test_padding_invariance_and_zero_outputs_past_length():
    B, T, I, H = 3, 7, 4, 5
    x, W, b = rnd(0, B, T, I), rnd(1, I + H, 4 * H) * 0.5, rnd(2, 4 * H) * 0.1
    lens = np.array([7, 3, 5])
    out, (c, h), _ = O.lstm_seq_fwd(x, lens, W, b)
    x2 = x.copy()
    x2[1, 3:] = 99.0  # garbage in the padding must not matter
    out2, (c2, h2), _ = O.lstm_seq_fwd(x2, lens, W, b)
    assert np.array_equal(out, out2) and np.array_equal(c, c2) and np.array_equal(h, h2)
    assert np.all(out[1, 3:] == 0) and np.all(out[2, 5:] == 0)
    assert np.allclose(h[1], out[1, 2]) and np.allclose(h[0], out[0, 6])  # final state = last valid step


def test_birnn_backward_direction_is_forward_on_reversed_input():
    B, T, I, H = 2, 6, 3, 4
    x = rnd(3, B, T, I)
    lens = np.array([6, 4])
    x *= (np.arange(T)[None, :, None] < lens[:, None, None])
    fw = [(rnd(4, I + H, 4 * H) * 0.4, rnd(5, 4 * H) * 0.1)]
    bw = [(rnd(6, I + H, 4 * H) * 0.4, rnd(7, 4 * H) * 0.1)]
    out, (sf, sb), _ = O.birnn_fwd(x, lens, fw, bw)
    ob, (cb, hb), _ = O.lstm_seq_fwd(O.reverse_sequence(x, lens), lens, *bw[0])
    assert np.allclose(out[..., H:], O.reverse_sequence(ob, lens))
    assert np.allclose(sb[0][1], hb)
    assert np.allclose(out[1, 0, H:], hb[1])  # bw final state sits at t = 0 of the un-reversed output


def test_concat_matmul_equals_split_matmul():
    B, I, H = 4, 6, 5
    x, h, c = rnd(8, B, I), rnd(9, B, H), rnd(10, B, H) * 0.3
    W, b = rnd(11, I + H, 4 * H), rnd(12, 4 * H)
    h1, c1, _ = O.lstm_cell(np.concatenate([x, h], 1), c, W, b)
    z = x @ W[:I] + h @ W[I:] + b
    i, j, f, o = np.split(z, 4, axis=1)
    c2 = np.clip(O.sigmoid(f + 1) * c + O.sigmoid(i) * np.tanh(j), -1, 1)
    assert np.allclose(c1, c2) and np.allclose(h1, O.sigmoid(o) * np.tanh(c2))


def test_attention_alignments_are_masked_distributions():
    B, T, Dx, H, Tm, Dm = 3, 4, 3, 6, 7, 5
    for kind in ('luong', 'scaled_luong', 'bahdanau', 'normed_bahdanau'):
        mem_len = np.array([7, 2, 5])
        spec = O.AttnSpec(kind=kind, memory=rnd(13, B, Tm, Dm), mem_len=mem_len, Wm=rnd(14, Dm, H), Wl=rnd(15, H + Dm, H),
                          Wq=rnd(16, H, H), v=rnd(17, H), g=np.asarray(0.8), b=rnd(18, H) * 0.1)
        r = O.attn_rnn_fwd(rnd(19, B, T, Dx), np.array([4, 4, 2]), rnd(20, Dx + 2 * H, 4 * H) * 0.3, np.zeros(4 * H),
                           [spec])
        a = r['alignments'][0]
        assert np.allclose(a[:2].sum(-1), 1.0) and np.allclose(a[2, :2].sum(-1), 1.0)
        assert np.all(a[1, :, 2:] == 0) and np.all(a[2, :, 5:] == 0)   # exactly zero past memory_len
        assert np.all(a[2, 2:] == 0) and np.all(r['outputs'][2, 2:] == 0)  # zero past the query length
        assert r['outputs'].shape[-1] == H and spec.output_attention == ('luong' in kind)


def test_sequence_loss_and_clip_adam_identities():
    B, T, V = 3, 5, 7
    logits, lens = rnd(21, B, T, V), np.array([5, 2, 4])
    targets = np.random.default_rng(22).integers(0, V, (B, T))
    loss, d = O.sequence_loss_fwd_bwd(logits, targets, lens)
    manual = 0.0
    for b in range(B):
        for t in range(lens[b]):
            z = logits[b, t]
            manual += np.log(np.exp(z).sum()) - z[targets[b, t]]
    assert np.isclose(loss, manual / lens.sum())
    assert np.all(d[1, 2:] == 0) and np.allclose(d.sum(-1), 0)
    P = {'w': np.ones(4)}; G = {'w': np.array([3.0, 4.0, 0.0, 0.0])}
    m, v = {'w': np.zeros(4)}, {'w': np.zeros(4)}
    gn = O.clip_and_adam(P, G, m, v, 0, 1e-3, clip=1.0, warmup_steps=750)
    assert np.isclose(gn, 5.0)
    lr = 1e-3 / 750
    g = np.array([0.6, 0.8, 0, 0])  # clipped to unit norm
    lr_t = lr * np.sqrt(1 - 0.999) / (1 - 0.9)
    assert np.allclose(P['w'], 1 - lr_t * (0.1 * g) / (np.sqrt(0.001 * g * g) + 1e-8))


def test_gather_tree_and_beam_shapes_on_tiny_model():
    from avsr_tf1_b200.seq2seq import Seq2SeqModel
    from tests.helpers import cast_batch, config_hparams, oracle_hparams, synthetic_batch, to_data_sequences
    hp = config_hparams(1, units=8, embedding_size=6, beam_width=3)
    hp.max_label_length = 7
    batch = synthetic_batch(hp, B=2, Ta=5, Fa=4, L=3)
    m = Seq2SeqModel(to_data_sequences(batch), 'train', hp, device='cpu')
    om = O.OracleModel(oracle_hparams(hp), {k: v.astype(np.float64) for k, v in m.store.to_numpy('p').items()})
    r = om.beam_decode(cast_batch(batch, np.float64))
    T = r['step_ids'].shape[1]
    assert r['predicted_ids'].shape == (2, T, 3) and T <= 7
    assert np.all(np.diff(r['scores'], axis=2) <= 1e-12)  # top_k returns beams best-first
    g = om.greedy_decode(cast_batch(batch, np.float64))
    assert g.shape[0] == 2 and g.shape[1] <= 7
    for row in g:  # after EOS everything is 0 (impute_finished)
        hits = np.nonzero(row == 29)[0]
        if hits.size:
            assert np.all(row[hits[0] + 1:] == 0)


def test_generator_known_answers_and_uniformity():
    """rand_u32 restates csrc/common.cuh avsr_rand_u32: the words below were printed by the C function compiled for
    the host (nvcc, same header).  Keep rates follow the thresholds; streams / steps / hi words decorrelate."""
    from oracle.avsr_oracle import DropSpec, keep_threshold, rand_u32
    assert rand_u32(1, 2, 3, 4, np.arange(4)).tolist() == [1982725394, 627806037, 120824804, 4077171900]
    assert int(rand_u32(0xDEADBEEF, 123456, 77, 4000000000, 4294967295)) == 873638011
    n = 200000
    w = rand_u32(7, 0, 9, 0, np.arange(n)).astype(np.float64) / 2.0 ** 32
    assert abs(w.mean() - 0.5) < 5e-3 and abs((w < 0.9).mean() - 0.9) < 3e-3
    for other in (rand_u32(7, 1, 9, 0, np.arange(n)), rand_u32(7, 0, 10, 0, np.arange(n)),
                  rand_u32(7, 0, 9, 1, np.arange(n)), rand_u32(8, 0, 9, 0, np.arange(n))):
        c = np.corrcoef(w, other.astype(np.float64))[0, 1]
        assert abs(c) < 0.01
    assert keep_threshold(1.0) == 0 and keep_threshold(0.5) == 2 ** 31
    spec = DropSpec((3, 4), 16, (0.9, 0.8, 1.0))
    fx = spec.x_factor(4, 50, 64, np.float64)
    assert set(np.unique(fx)) == {0.0, 4294967296.0 / keep_threshold(0.9)}
    assert abs(fx.mean() - 1.0) < 0.02  # inverted dropout is unbiased
    assert np.all(spec.step_factor(2, 5, 4, 64, np.float64) == 1.0)  # output keep 1.0: no mask


# ---- hand-computed known answers for the pieces no PyTorch primitive pins (cell_clip, the Bahdanau
# scorers, one beam-search step) -------------------------------------------------------------------------------------
def test_cell_clip_known_answer_and_zero_gradient_outside_the_range():
    """cell_clip = 1.0 (cells.py:16): c = clip(f*c_prev + i*j, -1, 1) with saturated gates i = j = f = o = 1;
    the gradient wrt c_prev is f inside the range and 0 where the clip is active."""
    big = 40.0
    W = np.zeros((2, 4))
    b = np.array([big, big, big, big])            # i, j, f (+1 forget bias), o all saturate at 1 (tanh(40) = 1)
    for c_prev, c_want in ((0.5, 1.0), (-0.5, 0.5), (-3.0, -1.0)):
        h, c, cache = O.lstm_cell(np.zeros((1, 2)), np.array([[c_prev]]), W, b)
        assert np.allclose(c, c_want) and np.allclose(h, np.tanh(c_want))
        dxh, dc_prev, _, _ = O.lstm_cell_bwd(np.zeros((1, 1)), np.ones((1, 1)), cache, W)
        inside = -1.0 <= c_prev + 1.0 <= 1.0
        assert np.allclose(dc_prev, 1.0 if inside else 0.0), (c_prev, dc_prev)


def _mech(kind, **kw):
    mem = np.array([[[1.0, 0.0], [0.0, 1.0]]])  # B = 1, Tm = 2, Dm = 2; memory_layer = identity -> keys = memory
    base = dict(kind=kind, memory=mem, mem_len=np.array([2]), Wm=np.eye(2), Wl=np.eye(4)[:, :2])
    base.update(kw)
    return O.AttnSpec(**base)


def test_bahdanau_scores_known_answer():
    """score_t = sum_u v_u tanh(keys_tu + (q Wq)_u) (attention.py:25-33)."""
    from math import tanh
    spec = _mech('bahdanau', Wq=np.eye(2), v=np.array([1.0, 2.0]))
    _, keys, _ = O._prepare_memory(spec)
    score, _ = O._score_fwd(spec, keys, np.array([[0.5, -0.5]]))
    want = [1.0 * tanh(1.0 + 0.5) + 2.0 * tanh(0.0 - 0.5), 1.0 * tanh(0.0 + 0.5) + 2.0 * tanh(1.0 - 0.5)]
    assert np.allclose(score, [want])


def test_normed_bahdanau_scores_known_answer():
    """v_hat = g v / |v|, bias inside the tanh (attention.py:34-42): v = (3, 4) -> |v| = 5."""
    from math import tanh
    spec = _mech('normed_bahdanau', Wq=np.eye(2), v=np.array([3.0, 4.0]), g=np.asarray(0.5), b=np.array([0.1, -0.2]))
    _, keys, _ = O._prepare_memory(spec)
    score, _ = O._score_fwd(spec, keys, np.array([[0.5, -0.5]]))
    nv = (0.5 * 3.0 / 5.0, 0.5 * 4.0 / 5.0)
    want = [nv[0] * tanh(1.0 + 0.5 + 0.1) + nv[1] * tanh(0.0 - 0.5 - 0.2),
            nv[0] * tanh(0.0 + 0.5 + 0.1) + nv[1] * tanh(1.0 - 0.5 - 0.2)]
    assert np.allclose(score, [want])


def test_scaled_luong_and_memory_mask_known_answer():
    """score = g keys.q (attention.py:64-72); rows past memory_sequence_length are zeroed BEFORE the key projection
    and their scores masked to -inf, so a memory of length 1 puts all the weight on row 0."""
    spec = _mech('scaled_luong', g=np.asarray(2.0))
    _, keys, _ = O._prepare_memory(spec)
    score, _ = O._score_fwd(spec, keys, np.array([[0.5, -0.25]]))
    assert np.allclose(score, [[1.0, -0.5]])
    short = _mech('luong', mem_len=np.array([1]))
    values, keys, mask = O._prepare_memory(short)
    assert np.array_equal(values[0, 1], [0.0, 0.0]) and np.array_equal(mask, [[1.0, 0.0]])
    x = np.zeros((1, 1, 1))
    W = np.zeros((1 + 2 + 2, 8))
    r = O.attn_rnn_fwd(x, np.array([1]), W, np.zeros(8), [short])
    assert np.allclose(r['alignments'][0][0, 0], [1.0, 0.0]) and np.allclose(r['contexts'][0][0, 0], [1.0, 0.0])


def test_beam_search_step_known_answer():
    """One BeamSearchDecoder step by hand (W = 2, V = 3, EOS = 2, length_penalty_weight = 0.6).
    Step 1 from the initial state (log_probs = [0, -inf]): p = (0.5, 0.3, 0.2) on beam 0.  Candidates: word 0 with
    score log 0.5 / ((5+1)/6)^0.6 = -0.693, word 1 with log 0.3 = -1.204, EOS (length stays 0) with
    log 0.2 / (5/6)^0.6 = -1.795: the survivors are words 0 and 1, both from parent 0."""
    eos, lpw = 2, 0.6
    logits = np.log(np.array([[[0.5, 0.3, 0.2], [0.1, 0.1, 0.8]]]))
    lp0 = np.array([[0.0, -np.inf]])
    word, parent, score, lp, fin, ln = O.beam_search_step(logits, lp0, np.zeros((1, 2), bool), np.zeros((1, 2), np.int64),
                                                           eos, lpw)
    assert word.tolist() == [[0, 1]] and parent.tolist() == [[0, 0]]
    assert np.allclose(score, [[np.log(0.5), np.log(0.3)]]) and np.allclose(lp, score)
    assert fin.tolist() == [[False, False]] and ln.tolist() == [[1, 1]]
    # Step 2: beam 0 (log p = log 0.5) sees p = (0.1, 0.1, 0.8); beam 1 (log 0.3) sees p = (0.6, 0.3, 0.1).
    #   beam 0 + EOS : (log 0.5 + log 0.8) / (6/6)^0.6          = -0.9163  (EOS does not lengthen: length 1)
    #   beam 1 + w0  : (log 0.3 + log 0.6) / (7/6)^0.6          = -1.5633
    #   beam 0 + w0/1: (log 0.5 + log 0.1) / (7/6)^0.6          = -2.7311
    logits2 = np.log(np.array([[[0.1, 0.1, 0.8], [0.6, 0.3, 0.1]]]))
    word, parent, score, lp, fin, ln = O.beam_search_step(logits2, lp, fin, ln, eos, lpw)
    assert word.tolist() == [[2, 0]] and parent.tolist() == [[0, 1]]
    assert np.allclose(score, [[np.log(0.4), np.log(0.18) / (7.0 / 6.0) ** 0.6]])
    assert np.allclose(lp, [[np.log(0.4), np.log(0.18)]])
    # state lengths count the EOS of a beam that finishes now (TF: "beams that are now finished have their length
    # increased by 1"), although the EOS candidate was SCORED with the un-lengthened hypothesis
    assert fin.tolist() == [[True, False]] and ln.tolist() == [[2, 2]]
    # Step 3: the finished beam may only be extended by EOS at zero cost and keeps its length
    logits3 = np.log(np.array([[[0.98, 0.01, 0.01], [0.05, 0.05, 0.9]]]))
    word, parent, score, lp, fin, ln = O.beam_search_step(logits3, lp, fin, ln, eos, lpw)
    #   finished beam 0 + EOS: log 0.4 / (7/6)^0.6 = -0.835 (its length 2 now includes the EOS);
    #   beam 1 + EOS: (log 0.18 + log 0.9) / (7/6)^0.6 = -1.659
    assert word.tolist() == [[2, 2]] and parent.tolist() == [[0, 1]]
    assert np.allclose(score, [[np.log(0.4) / (7.0 / 6.0) ** 0.6, np.log(0.162) / (7.0 / 6.0) ** 0.6]])
    assert np.allclose(lp, [[np.log(0.4), np.log(0.162)]])
    assert fin.tolist() == [[True, True]] and ln.tolist() == [[2, 3]]
    # ties: equal scores keep the lowest flat index first (tf.nn.top_k)
    tie = np.log(np.array([[[0.25, 0.25, 0.5], [0.25, 0.25, 0.5]]]))
    word, parent, *_ = O.beam_search_step(tie, np.array([[0.0, 0.0]]), np.zeros((1, 2), bool), np.ones((1, 2), np.int64), eos, 0.0)
    assert word.tolist() == [[2, 2]] and parent.tolist() == [[0, 1]]


def test_devel_losses_known_answers():
    """devel.py:12-52 on uniform logits over 3 classes (p = 1/3 each), label 0, worked by hand:
    mc_loss = -ln(1/3) - 2 ln(2/3); focal_loss (gamma 2) = -(2/3)^2 ln(1/3) - 2 (1/3)^2 ln(2/3)."""
    z = np.zeros((1, 1, 3))
    y = np.zeros((1, 1), np.int64)
    lens = np.array([1])
    mc, dmc = O.sequence_loss_fwd_bwd(z, y, lens, loss_fun='mc_loss')
    fo, dfo = O.sequence_loss_fwd_bwd(z, y, lens, loss_fun='focal_loss')
    # (sequence_loss divides by token count + 1e-12)
    assert abs(mc - (-np.log(1 / 3) - 2 * np.log(2 / 3))) < 1e-10
    assert abs(fo - (-(2 / 3) ** 2 * np.log(1 / 3) - 2 * (1 / 3) ** 2 * np.log(2 / 3))) < 1e-10
    # softmax Jacobian at the uniform point: dz_j = p (g_j - mean(g)), g = dL/dp; rows of a softmax gradient sum to zero
    g = np.array([-3.0, 1.5, 1.5])  # mc_loss: -1/p for the label, 1/(1-p) otherwise
    np.testing.assert_allclose(dmc[0, 0], (g - g.mean()) / 3, atol=1e-10)
    assert abs(dfo.sum()) < 1e-12 and dfo[0, 0, 0] < 0 < dfo[0, 0, 1]
    # past the label length nothing contributes
    z2 = np.zeros((1, 2, 3))
    mc2, d2 = O.sequence_loss_fwd_bwd(z2, np.zeros((1, 2), np.int64), lens, loss_fun='mc_loss')
    assert abs(mc2 - mc) < 1e-10 and np.all(d2[0, 1] == 0.0)


def test_highway_and_instance_norm_known_answers():
    """HighwayWrapper (cells.py:89-90; zero carry kernel, carry bias 1 -> carry = sigmoid(1) everywhere) around a cell
    whose weights are zero (h = o tanh(c) = 0): the output is the carried input alone.  instance_norm: a feature that
    runs 0, 1, 2 along time normalises to -sqrt(1.5), 0, sqrt(1.5) (variance 2/3, epsilon 1e-6)."""
    B, T, H = 1, 3, 2
    x = np.array([[[1.0, -2.0], [0.5, 0.25], [3.0, 0.0]]])
    zero_layer = (np.zeros((2 * H, 4 * H)), np.zeros(4 * H))
    out, _, _ = O.stacked_lstm_fwd(x, np.array([T]), [zero_layer, zero_layer],
                                   highway=[None, (np.zeros((H, H)), np.ones(H))])
    # layer 0 emits zeros, so layer 1 carries zeros: feed x straight into the wrapped layer instead
    out1, _, _ = O.stacked_lstm_fwd(x, np.array([T]), [zero_layer], highway=[(np.zeros((H, H)), np.ones(H))])
    carry = 1.0 / (1.0 + np.exp(-1.0))
    np.testing.assert_allclose(out, 0.0, atol=1e-15)
    np.testing.assert_allclose(out1, x * carry, atol=1e-12)
    hp = O.OracleHParams(batch_normalisation=False, instance_normalisation=True)
    P = {'audio/InstanceNorm/gamma': np.array([2.0]), 'audio/InstanceNorm/beta': np.array([0.5])}
    om = O.OracleModel(hp, P)
    y = om._bn('audio', np.array([[[0.0], [1.0], [2.0]]]), True, {})
    ref = np.array([-1.0, 0.0, 1.0]) / np.sqrt(2.0 / 3.0 + 1e-6) * 2.0 + 0.5
    np.testing.assert_allclose(y[0, :, 0], ref, atol=1e-12)
