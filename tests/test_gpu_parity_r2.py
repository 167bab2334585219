"""Round-2 parity tests that close the gaps the round-1 review named:
  * the BENCHMARKED shape itself (B = 256, Ta = 300, Tv = 75 lip crops, 41 label steps) is checked: tensor-core mode
    (persistent tcgen05 kernels) against the exact-fp32 mode of the same library (per-step CUDA-core kernels, which the
    small-shape tests pin to the oracle at 2e-5), on the parity graph and with every DropoutWrapper on;
  * beam search at the reference's default width 10 with realistic (unsharpened) logits;
  * a 200-step training run in tensor-core mode tracks the exact-fp32 run (justifies the gradient tolerance of the
    tensor-core tests).
"""
import numpy as np
import pytest
import torch

from oracle import avsr_oracle as O
from tests.helpers import cast_batch, config_hparams, oracle_hparams, synthetic_batch, to_data_sequences

pytestmark = pytest.mark.gpu


def scaled_err(got, want):
    got, want = got.double(), want.double()
    return float((got - want).abs().max() / want.abs().max().clamp_min(1e-30))


def run_once(hp, ds, tc, seed=2001, step_word=7):
    from avsr_tf1_b200 import ops
    from avsr_tf1_b200.seq2seq import Seq2SeqModel
    old = ops.set_tensor_cores(tc)
    try:
        model = Seq2SeqModel(ds, 'train', hp, seed=seed)
        model._global_step = step_word
        model.feed(ds)
        model._set_step_scalars()
        model.forward_backward()
        model.finish_gradients()
        loss, gnorm = model.fetch_scalars()
        probes = {}
        for key, enc in (('video', model._video_encoder), ('audio', model._audio_encoder)):
            if enc is None:
                continue
            d = enc.get_data()
            probes[key + '/outputs'] = d.outputs.clone()
            probes[key + '/final_c'] = d.final_state[0].clone()
            probes[key + '/final_h'] = d.final_state[1].clone()
        # contexts of steps past an utterance's length are don't-care values (the persistent kernels skip the sweep,
        # the per-step kernels compute it; nothing reads them): compare the valid steps
        def valid(lens, T):
            return (torch.arange(T, device='cuda')[:, None] < lens.cuda()[None, :]).float()[:, :, None]
        if hasattr(model._audio_encoder, 'attention_contexts'):  # AV-Align
            Ha = model._audio_encoder._num_units_per_layer[-1]
            ctx_a = model._audio_encoder.attention_contexts[:, :, Ha:]
            probes['audio/xmodal_contexts'] = ctx_a * valid(model._in['audio_len'], ctx_a.shape[0])
        H = model._decoder._H
        for k, mb in enumerate(model._decoder._cell.bufs):  # (two mechanisms for the WLAS decoder)
            ctx_d = mb.hc[:, :, H:]
            probes['decoder/contexts' + ('_%d' % k if k else '')] = ctx_d * valid(model._in['labels_len'], ctx_d.shape[0])
        probes['decoder/logits'] = model._decoder._logits.clone()
        grads = {k: torch.from_numpy(v) for k, v in model.store.to_numpy('g').items()}
        launches = None
        return loss, gnorm, probes, grads, launches
    finally:
        ops.set_tensor_cores(old)


@pytest.mark.parametrize('graph', ['parity', 'dropout'])
def test_bench_shape_tensor_core_mode_tracks_exact_fp32(graph):
    """B = 256 x Ta = 300 x Tv = 75 x 3888-d crops x 41 label steps: 32 clusters x 8 utterances, full lengths - the shape
    bench.py times.  1e-3 on encoder states and attention contexts (north_star), loss 1e-3, gradients 1.5e-2."""
    over = dict(use_dropout=True) if graph == 'dropout' else {}
    hp = config_hparams(5, **over)
    batch = synthetic_batch(hp, B=256, Ta=300, Tv=75, Fa=80, Fv=3888, L=40, ragged=False)
    ds = to_data_sequences(batch)
    loss_x, gn_x, px, gx, _ = run_once(hp, ds, tc=False)
    loss_t, gn_t, pt, gt, _ = run_once(hp, ds, tc=True)
    assert abs(loss_t - loss_x) <= 1e-3 * abs(loss_x), (loss_t, loss_x)
    worst = {}
    for k in px:
        worst[k] = scaled_err(pt[k], px[k])
    print('bench-shape scaled errors (tensor-core vs exact fp32):', {k: '%.2e' % v for k, v in worst.items()})
    for k, v in worst.items():
        assert v <= 1e-3, f'{k}: scaled error {v:.3e}'
    assert abs(gn_t - gn_x) <= 5e-3 * gn_x, (gn_t, gn_x)
    gmax = max(float(g.abs().max()) for g in gx.values())
    for k in gx:
        scale = max(float(gx[k].abs().max()), 1e-3 * gmax)
        err = float((gt[k].double() - gx[k].double()).abs().max()) / scale
        assert err <= 1.5e-2, f'{k}: gradient scaled error {err:.3e}'


@pytest.mark.parametrize('cfg,B', [(4, 128), (2, 64)])
def test_bench_shapes_of_the_other_configurations(cfg, B):
    """BASELINE configs 4 (WLAS: the dual-attention cluster-of-8 kernels, 8 clusters x 16 utterances) and 2 (BiLSTM +
    Bahdanau: the two-product Bahdanau kernels over a 512-deep memory) at the batch and lengths bench.py times, reference
    default graph (every DropoutWrapper on): tensor-core mode against the exact-fp32 mode, same bars as above."""
    hp = config_hparams(cfg, use_dropout=True)
    batch = synthetic_batch(hp, B=B, Ta=300, Tv=75, Fa=80, Fv=128, L=40, ragged=False)
    ds = to_data_sequences(batch)
    loss_x, gn_x, px, gx, _ = run_once(hp, ds, tc=False)
    loss_t, gn_t, pt, gt, _ = run_once(hp, ds, tc=True)
    assert abs(loss_t - loss_x) <= 1e-3 * abs(loss_x), (loss_t, loss_x)
    worst = {k: scaled_err(pt[k], px[k]) for k in px}
    print('config %d bench-shape scaled errors (tensor-core vs exact fp32):' % cfg, {k: '%.2e' % v for k, v in worst.items()})
    for k, v in worst.items():
        assert v <= 1e-3, f'{k}: scaled error {v:.3e}'
    assert abs(gn_t - gn_x) <= 5e-3 * gn_x, (gn_t, gn_x)
    gmax = max(float(g.abs().max()) for g in gx.values())
    for k in gx:
        scale = max(float(gx[k].abs().max()), 1e-3 * gmax)
        err = float((gt[k].double() - gx[k].double()).abs().max()) / scale
        assert err <= 1.5e-2, f'{k}: gradient scaled error {err:.3e}'


def test_bench_shape_ragged_lengths_tensor_core_mode_tracks_exact_fp32():
    """Same batch size with ragged lengths (every mask of the persistent kernels at 32 clusters)."""
    hp = config_hparams(5, use_dropout=True)
    batch = synthetic_batch(hp, B=250, Ta=120, Tv=40, Fa=80, Fv=128, L=20, ragged=True)
    ds = to_data_sequences(batch)
    loss_x, gn_x, px, gx, _ = run_once(hp, ds, tc=False)
    loss_t, gn_t, pt, gt, _ = run_once(hp, ds, tc=True)
    assert abs(loss_t - loss_x) <= 1e-3 * abs(loss_x), (loss_t, loss_x)
    for k in px:
        err = scaled_err(pt[k], px[k])
        assert err <= 1e-3, f'{k}: scaled error {err:.3e}'
    assert abs(gn_t - gn_x) <= 5e-3 * gn_x


def test_beam_search_default_width_realistic_logits():
    """beam_width = 10 (avsr.py:59), B = 16, NO sharpening of the output layer: the logits are those of a model two
    training steps from its initialisation, i.e. nearly flat - the hardest case for top-k agreement.  Exact-fp32 mode
    against the oracle in fp32: step ids / parents of every beam must agree except where two candidates are within
    float rounding of each other; such near-ties are counted and bounded, and the scores agree to 1e-4."""
    from avsr_tf1_b200 import ops, utils
    from avsr_tf1_b200.seq2seq import Seq2SeqModel
    old = ops.set_tensor_cores(False)
    try:
        hp = config_hparams(5, decoding_algorithm='beam_search', beam_width=10)
        hp.max_label_length = 14
        batch = synthetic_batch(hp, B=16, Ta=30, Tv=10, L=6, ragged=True)
        ds = to_data_sequences(batch)
        train = Seq2SeqModel(ds, 'train', hp, seed=2001)
        for _ in range(2):
            train.train_step(ds)
        ev = Seq2SeqModel(ds, 'evaluate', hp, share_params_with=train)
        ids = ev.predict(ds)
        P = {k: v.astype(np.float32) for k, v in train.store.to_numpy('p').items()}
        r = O.OracleModel(oracle_hparams(hp), P).beam_decode(cast_batch(batch, np.float32))
        ref = r['predicted_ids'][:, :, 0]
        bo = ev._decoder.beam_search_output
        assert bo.scores.shape == r['scores'].shape, (bo.scores.shape, r['scores'].shape)
        # scores of the surviving beams are sorted, so they are comparable even where near-ties permute the beams
        np.testing.assert_allclose(bo.scores, r['scores'], rtol=1e-4, atol=1e-4)
        step_mismatch = (bo.predicted_ids != r['step_ids']) | (bo.parent_ids != r['parent_ids'])
        flipped_utts = int((ids != ref).any(axis=1).sum()) if ids.shape == ref.shape else ids.shape[0]
        print('beam width 10: %d of %d (utterance, step, beam) entries differ, %d of %d best hypotheses differ'
              % (int(step_mismatch.sum()), step_mismatch.size, flipped_utts, ids.shape[0]))
        # where entries differ, the two candidates' scores are within float rounding (a genuine near-tie)
        if step_mismatch.any():
            b, t, w = np.nonzero(step_mismatch)
            gap = np.abs(bo.scores[b, t, w] - r['scores'][b, t, w])
            assert gap.max() <= 1e-4, gap.max()
        assert flipped_utts <= 1, flipped_utts
        ud = hp.unit_dict
        pred = {f'utt{b}': utils.ids_to_symbols(ids[b], ud) for b in range(ids.shape[0])}
        truth = {f'utt{b}': utils.ids_to_symbols(batch['labels'][b], ud) for b in range(ids.shape[0])}
        assert utils.compute_wer(pred, truth) == O.compute_wer(pred, truth)
    finally:
        ops.set_tensor_cores(old)


def test_two_hundred_steps_tensor_core_training_tracks_exact_fp32():
    """200 Adam steps on a rotating set of batches.  Two optimisation runs that differ by rounding drift apart
    chaotically while the loss falls fast, so the test separates the two questions:
      (a) per-step fidelity ALONG a real training trajectory: every 10th step the parameters of the exact-fp32 run are
          copied into a tensor-core-mode model and loss / gradient of the same batch are compared (loss 1e-3, global
          norm 5e-3, gradient direction cosine >= 0.998; the trajectory itself is not reproducible - fp32 atomics order - and
          the smallest cosine of a run has been seen between 0.99905 and 0.99995) - this is what the 1.5e-2 per-tensor gradient tolerance of the
          tensor-core tests has to guarantee;
      (b) the independent tensor-core run learns the same thing: same final loss level (within the band two chaotic
          trajectories oscillate in) and step-by-step agreement over the first 10 steps."""
    from avsr_tf1_b200 import ops
    from avsr_tf1_b200.seq2seq import Seq2SeqModel
    hp = config_hparams(5, learning_rate=1e-3)
    hp.kwargs['warmup_steps'] = 20
    batches = [to_data_sequences(synthetic_batch(hp, B=8, Ta=40, Tv=12, L=8, ragged=True, seed=s)) for s in range(4)]
    old = ops.set_tensor_cores(False)
    try:
        exact = Seq2SeqModel(batches[0], 'train', hp, seed=2001)
        probe = Seq2SeqModel(batches[0], 'train', hp, seed=2001)
        curve_x, worst = [], dict(loss=0.0, gnorm=0.0, cos=1.0)
        for s in range(200):
            ds = batches[s % 4]
            if s % 10 == 0:
                ops.set_tensor_cores(True)
                probe.store.flat.copy_(exact.store.flat)
                probe.store.sync_tf32()
                probe._global_step = exact._global_step
                probe.feed(ds)
                probe._set_step_scalars()
                probe.forward_backward()
                probe.finish_gradients()
                loss_t, gn_t = probe.fetch_scalars()
                g_t = probe.store.grad.double().clone()
                ops.set_tensor_cores(False)
            loss_x, gn_x = exact.train_step(ds)
            curve_x.append(loss_x)
            if s % 10 == 0:
                g_x = exact.store.grad.double()
                cos = float((g_t * g_x).sum() / (g_t.norm() * g_x.norm()))
                worst['loss'] = max(worst['loss'], abs(loss_t - loss_x) / abs(loss_x))
                worst['gnorm'] = max(worst['gnorm'], abs(gn_t - gn_x) / gn_x)
                worst['cos'] = min(worst['cos'], cos)
        ops.set_tensor_cores(True)
        tcm = Seq2SeqModel(batches[0], 'train', hp, seed=2001)
        tcm.use_cuda_graph = True  # (four batch shapes, captured once each)
        curve_t = [tcm.train_step(batches[s % 4])[0] for s in range(200)]
    finally:
        ops.set_tensor_cores(old)
    x, t = np.array(curve_x), np.array(curve_t)
    print('200 steps: loss %.4f -> %.4f (exact fp32) / %.4f (tensor cores, independent run); along the exact trajectory: '
          'loss gap %.2e, global-norm gap %.2e, min gradient cosine %.6f'
          % (x[0], x[-20:].mean(), t[-20:].mean(), worst['loss'], worst['gnorm'], worst['cos']))
    assert np.isfinite(t).all() and x[-20:].mean() < 0.5 * x[:4].mean()  # the model does learn over the run
    # (the smallest cosine of a run moves with the trajectory, which fp32 atomics make different every time: 0.99905 ..
    # 0.99995 over a dozen runs, the low values late in training where the gradient is a small difference of large terms)
    assert worst['loss'] <= 1e-3 and worst['gnorm'] <= 5e-3 and worst['cos'] >= 0.998, worst
    # independent runs drift apart (the loss oscillates between 0.24 and 0.35 at lr 1e-3 on four memorised batches):
    # same level, not the same value
    # (the last-20-step means of two such runs have been seen anywhere in 0.24 .. 0.40: a factor 2 bounds the band)
    lo, hi = sorted([float(t[-20:].mean()), float(x[-20:].mean())])
    assert hi <= 2.0 * lo and hi < 0.15 * x[:4].mean()
    assert np.abs(t[:10] - x[:10]).max() <= 1e-3 * x[:10].max()  # before the runs drift apart they agree step by step
