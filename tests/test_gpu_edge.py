"""Edge cases of the path through the C ABI against the oracle: one-utterance batches, one-step sequences, utterances of
a single frame inside a ragged batch, a memory longer than the persistent attention kernel holds (it must hand over to
the step-wise kernels), the decoding cap."""
import numpy as np
import pytest

from oracle import avsr_oracle as O
from tests.helpers import cast_batch, config_hparams, oracle_hparams, synthetic_batch, to_data_sequences
from tests.test_gpu_model import tensor_cores  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu


def run(cfg, batch, over=None):
    from avsr_tf1_b200.seq2seq import Seq2SeqModel
    hp = config_hparams(cfg, **(over or {}))
    ds = to_data_sequences(batch)
    model = Seq2SeqModel(ds, 'train', hp, seed=2001)
    P = {k: v.astype(np.float64) for k, v in model.store.to_numpy('p').items()}
    loss_ref, G_ref, rec = O.OracleModel(oracle_hparams(hp), P).loss_and_grads(cast_batch(batch, np.float64))
    model.feed(ds)
    model._set_step_scalars()
    model.forward_backward()
    model.finish_gradients()
    loss, gnorm = model.fetch_scalars()
    return hp, model, loss, gnorm, loss_ref, G_ref


def check(loss, gnorm, loss_ref, G_ref):
    assert np.isfinite(loss) and np.isfinite(gnorm)
    assert abs(loss - loss_ref) <= 1e-3 * abs(loss_ref), (loss, loss_ref)
    gn_ref = O.global_norm(G_ref)
    assert abs(gnorm - gn_ref) <= 5e-3 * gn_ref, (gnorm, gn_ref)


@pytest.mark.parametrize('cfg', [1, 2, 4, 5])
@pytest.mark.parametrize('B,Ta,Tv,L', [(1, 1, 1, 1), (1, 9, 3, 2), (2, 1, 1, 1)])
def test_smallest_batches_and_sequences(cfg, B, Ta, Tv, L, tensor_cores):
    hp = config_hparams(cfg)
    batch = synthetic_batch(hp, B=B, Ta=Ta, Tv=Tv, L=L, ragged=False)
    _, _, loss, gnorm, loss_ref, G_ref = run(cfg, batch)
    check(loss, gnorm, loss_ref, G_ref)


@pytest.mark.parametrize('cfg', [2, 5])
def test_single_frame_utterances_inside_a_ragged_batch(cfg, tensor_cores):
    hp = config_hparams(cfg)
    batch = synthetic_batch(hp, B=6, Ta=24, Tv=8, L=6, ragged=True)
    for key in ('audio', 'video'):
        if key in batch:
            batch[key + '_len'][1] = 1
            batch[key + '_len'][4] = 1
            batch[key][1, 1:] = 0
            batch[key][4, 1:] = 0
    batch['labels_len'][2] = 1  # a target that is only EOS
    batch['labels'][2, 0] = 29
    batch['labels'][2, 1:] = 0
    _, _, loss, gnorm, loss_ref, G_ref = run(cfg, batch)
    check(loss, gnorm, loss_ref, G_ref)


def test_memory_longer_than_the_persistent_kernel_holds():
    """Decoder memory of 400 audio frames (> 384 rows the cluster kernel keeps scores for): the call must fall back to the
    step-wise attention kernels, with the same results."""
    from avsr_tf1_b200 import ops
    old = ops.set_tensor_cores(True)
    try:
        hp = config_hparams(1, units=256)
        batch = synthetic_batch(hp, B=3, Ta=400, Tv=8, L=5, ragged=True)
        batch['audio_len'][0] = 400
        ops.kernel_timing(True)
        _, _, loss, gnorm, loss_ref, G_ref = run(1, batch, dict(encoder_units_per_layer=((256,), (256,)),
                                                                decoder_units_per_layer=(256,)))
        times = ops.kernel_times()
        ops.kernel_timing(False)
        assert times['attn_lstm_fwd'][1] == 0 and times['attn_lstm_bwd'][1] == 0, times
        check(loss, gnorm, loss_ref, G_ref)
    finally:
        ops.set_tensor_cores(old)


def test_decoding_stops_at_max_label_length():
    """maximum_iterations = max_label_length (avsr.py:157, decoder_unimodal.py:214): a model that never emits EOS
    produces exactly that many ids; beam search likewise."""
    from avsr_tf1_b200.seq2seq import Seq2SeqModel
    for algo in ('greedy', 'beam_search'):
        hp = config_hparams(1, decoding_algorithm=algo, beam_width=3)
        hp.max_label_length = 12
        batch = synthetic_batch(hp, B=3, Ta=10, Tv=4, L=4, ragged=True)
        ds = to_data_sequences(batch)
        model = Seq2SeqModel(ds, 'evaluate', hp, seed=2001)
        bias = model.store.p('Decoder/decoder/my_dense/bias')
        bias[29] = -1e4  # EOS can never win
        model.store.sync_tf32()
        ids = model.predict(ds)
        assert ids.shape == (3, 12), ids.shape
        assert (ids != 29).all()


def test_small_input_bn_gamma_switches_to_the_explicit_gradient_path():
    """dgamma of the input normalisation is normally read off the layer-0 weight gradient DIVIDED by gamma
    (avsr_bn_input_grads); a gamma near zero would amplify the tf32 rounding of that gradient.  Seq2SeqModel checks the
    smallest |gamma| every BN_GAMMA_CHECK_EVERY steps and moves the encoder to the explicit path; the gradient of the tiny
    gamma must then agree with the oracle like every other entry."""
    import torch
    from avsr_tf1_b200.seq2seq import Seq2SeqModel
    hp = config_hparams(5)
    batch = synthetic_batch(hp, B=4, Ta=30, Tv=10, L=6, ragged=True)
    ds = to_data_sequences(batch)
    model = Seq2SeqModel(ds, 'train', hp, seed=2001)
    g = model.store.p('audio/batch_normalization/gamma')
    g[3] = 1e-6
    g[7] = -2e-3
    model.store.sync_tf32()
    assert not getattr(model._audio_encoder, 'explicit_bn_backward', False)
    model.feed(ds)
    model._set_step_scalars()
    model._guard_bn_shortcut()
    assert model._audio_encoder.explicit_bn_backward and not getattr(model._video_encoder, 'explicit_bn_backward', False)
    P = {k: v.astype(np.float64) for k, v in model.store.to_numpy('p').items()}
    loss_ref, G_ref, _ = O.OracleModel(oracle_hparams(hp), P).loss_and_grads(cast_batch(batch, np.float64))
    model.forward_backward()
    model.finish_gradients()
    loss, gnorm = model.fetch_scalars()
    check(loss, gnorm, loss_ref, G_ref)
    got = model.store.to_numpy('g')['audio/batch_normalization/gamma'].astype(np.float64)
    ref = G_ref['audio/batch_normalization/gamma']
    assert np.abs(got - ref).max() <= 2e-2 * np.abs(ref).max(), (got[[3, 7]], ref[[3, 7]])
    # and training goes on (graphs are re-captured on the new path)
    model.use_cuda_graph = True
    for _ in range(3):
        l, _ = model.train_step(ds)
        assert np.isfinite(l)
