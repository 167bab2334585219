"""Validates the oracle's analytic backward with central finite differences in fp64
(the restatement must be validated by independent means)."""
import numpy as np
import pytest

from avsr_tf1_b200.seq2seq import Seq2SeqModel
from oracle.avsr_oracle import OracleModel
from tests.helpers import add_aus, to_image_sequences, cast_batch, config_hparams, oracle_hparams, synthetic_batch, to_data_sequences

CASES = [
    (1, {}),
    (1, dict(attention_type=(('luong',), ('luong',)))),
    (1, dict(attention_type=(('bahdanau',), ('normed_bahdanau',)))),
    (2, {}),
    (3, {}),
    (4, {}),
    (4, dict(attention_type=(('bahdanau',), ('bahdanau',)))),
    (5, {}),
    (5, dict(attention_type=(('bahdanau',), ('scaled_luong',)))),
    (5, dict(batch_normalisation=False)),
]
# DropoutWrapper masks (fixed by the counter-based generator, so the loss stays differentiable) and decoder inputs
# chosen by scheduled sampling (the draws of a first pass are fed back: no gradient flows through them)
DROP = dict(use_dropout=True, audio_encoder_dropout_probability=(0.8, 0.7, 0.9),
            video_encoder_dropout_probability=(0.9, 0.8, 0.7), decoder_dropout_probability=(0.7, 0.9, 0.8))
CASES += [
    (1, DROP), (2, DROP), (4, DROP), (5, DROP),
    (5, dict(DROP, attention_type=(('bahdanau',), ('bahdanau',)))),
    (1, dict(sampling_probability_outputs=0.5)),
    (5, dict(DROP, sampling_probability_outputs=0.5)),
    (3, dict(regress_aus=True)), (4, dict(regress_aus=True)), (5, dict(DROP, regress_aus=True, au_loss_weight=3.0)),
    # resnet_cnn front-end on 8x8x3 crops (8 -> 4 -> 2 -> 1), through the encoders and the decoder
    (3, dict(video_processing='resnet_cnn', cnn_filters=(2, 3, 4, 5), cnn_dense_units=6)),
    (5, dict(DROP, video_processing='resnet_cnn', cnn_filters=(2, 3, 4, 5), cnn_dense_units=6, regress_aus=True)),
    # input_dense_layers (encoder.py:148-171): selu Dense stack between the input normalisation and the first layer -
    # uni / bidirectional (the state projections are named after it) / AV-Align, and with a CNN (its L2 term joins the loss)
    (1, dict(input_dense_layers=(5, 4))), (2, dict(input_dense_layers=(5,))), (5, dict(DROP, input_dense_layers=(4, 5))),
    (4, dict(input_dense_layers=(5,), video_processing='resnet_cnn', cnn_filters=(2, 3, 4, 5), cnn_dense_units=6)),
    # ResidualWrapper on encoder layers > 0 and the shared cell of layers 2.. (cells.py:77-92)
    (3, dict(DROP, residual_encoder=True)), (4, dict(residual_encoder=True, encoder_weight_sharing=True)),
    (3, dict(encoder_weight_sharing=True)),
    (1, dict(label_smoothing=0.1)), (5, dict(DROP, label_smoothing=0.2)),  # seq2seq.py:147-155
    # bimodal decoder with one stream missing (decoder_bimodal.py:127-142): zero state in the shared projection
    (4, dict(video_processing=None)), (4, dict(DROP, audio_processing=None)),
    (4, dict(audio_processing=None, encoder_units_per_layer=((5, 6, 6), (6, 6, 6)))),
    # HighwayWrapper on encoder layers > 0 (cells.py:89-90), alone, over the residual flag, with the shared cell
    (3, dict(DROP, highway_encoder=True)), (4, dict(highway_encoder=True, residual_encoder=True)),
    (3, dict(highway_encoder=True, encoder_weight_sharing=True)),
    # enable_attention=False (decoder_unimodal.py:319-327): the bare decoder cell started from the encoder state
    (1, dict(enable_attention=False)), (4, dict(DROP, enable_attention=False, sampling_probability_outputs=0.5)),
    (5, dict(enable_attention=False)),
    # instance_norm on the (batch-normalised) inputs (encoder.py:51-55)
    (1, dict(instance_normalisation=True)), (5, dict(DROP, instance_normalisation=True, batch_normalisation=False)),
    (2, dict(instance_normalisation=True, input_dense_layers=(5,))),
    # devel.py losses under sequence_loss (seq2seq.py:156-163)
    (1, dict(loss_fun='mc_loss')), (5, dict(DROP, loss_fun='focal_loss')),
    # one-hot decoder inputs (decoder_unimodal.py:75-76: embedding_size <= 0 -> tf.eye)
    (1, dict(embedding_size=0)), (5, dict(DROP, embedding_size=-1, sampling_probability_outputs=0.5)),
]


def tiny_model(cfg, over, seed=7):
    hp = config_hparams(cfg, units=6, **dict(dict(embedding_size=5), **over))
    batch = synthetic_batch(hp, B=3, Ta=7, Tv=5, Fa=4, Fv=3, L=4, ragged=True, seed=seed)
    if hp.regress_aus:
        add_aus(batch)
    if hp.video_processing == 'resnet_cnn':
        to_image_sequences(batch, hw=8)
    model = Seq2SeqModel(to_data_sequences(batch), 'train', hp, seed=11, device='cpu')
    P = {k: v.astype(np.float64) for k, v in model.store.to_numpy('p').items()}
    rng = np.random.default_rng(5)
    for k in P:  # break symmetric / zero initial values so every path carries gradient
        if k.startswith('CNN/') and k.endswith('kernel'):
            continue  # (He-initialised already; tripling them saturates the tiny network)
        if k.endswith(('bias', 'beta', 'attention_b')):
            P[k] = P[k] + 0.1 * rng.standard_normal(P[k].shape)
        elif k.endswith('kernel') or k.endswith('attention_v') or k.endswith('embedding_matrix'):
            P[k] = P[k] * 3.0
    return hp, cast_batch(batch, np.float64), P, model


@pytest.mark.parametrize('cfg,over', CASES)
def test_backward_matches_finite_differences(cfg, over):
    hp, batch, P, model = tiny_model(cfg, over)
    model._global_step = 2
    om = OracleModel(oracle_hparams(hp, model), P)
    if hp.sampling_probability_outputs > 0.0:
        _, rec = om.forward_train(batch)
        assert (rec['sample_ids'] >= 0).any()
        batch['dec_in_ids'] = rec['dec_in_ids']
    loss, G, _ = om.loss_and_grads(batch)
    assert np.isfinite(loss)
    trainable = set(model.store.names())
    assert set(G) == trainable
    rng = np.random.default_rng(3)
    eps = 1e-6
    worst = 0.0
    for name in sorted(trainable):
        flat = P[name].reshape(-1)
        for idx in rng.choice(flat.size, size=min(4, flat.size), replace=False):
            old = flat[idx]
            flat[idx] = old + eps
            lp, _ = om.forward_train(batch)
            flat[idx] = old - eps
            lm, _ = om.forward_train(batch)
            flat[idx] = old
            fd = (lp - lm) / (2 * eps)
            an = G[name].reshape(-1)[idx]
            err = abs(fd - an) / max(1e-6, abs(fd) + abs(an))
            # (gradients at rounding level - e.g. a conv bias in front of a batch norm, exactly 0 in theory - do not count)
            worst = max(worst, err if max(abs(fd), abs(an)) > 1e-7 else 0.0)
            assert abs(fd - an) <= 1e-6 + 2e-5 * max(abs(fd), abs(an)), (name, idx, fd, an)
    assert worst < 1e-3
