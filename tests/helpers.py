"""Shared test helpers: reference configs (BASELINE.json), synthetic batches
(SURVEY.md section 8d seeds) and product <-> oracle plumbing."""
from __future__ import annotations

import numpy as np

from avsr_tf1_b200 import make_batched_data, make_hparams

PARITY = dict(use_dropout=False, sampling_probability_outputs=0.0, regress_aus=False)


def config_hparams(cfg: int, units=None, **over):
    """hparams of BASELINE.json configs 1..5 with parity switches."""
    u3 = (units,) * 3 if units else (256,) * 3
    kw = dict(PARITY)
    if cfg == 1:
        u = units or 128
        kw.update(architecture='unimodal', audio_processing='features', encoder_units_per_layer=((u,), (u,)),
                  decoder_units_per_layer=(u,), batch_size=(2, 2))
    elif cfg == 2:
        kw.update(architecture='unimodal', audio_processing='features', encoder_type='bidirectional',
                  encoder_units_per_layer=(u3, u3), decoder_units_per_layer=(u3[0],),
                  attention_type=(('bahdanau',), ('bahdanau',)), batch_size=(64, 64))
    elif cfg == 3:
        kw.update(architecture='unimodal', video_processing='features', encoder_units_per_layer=(u3, u3),
                  decoder_units_per_layer=(u3[0],), batch_size=(64, 64))
    elif cfg == 4:
        kw.update(architecture='bimodal', video_processing='features', audio_processing='features',
                  encoder_units_per_layer=(u3, u3), decoder_units_per_layer=(u3[0],), batch_size=(128, 128))
    elif cfg == 5:
        kw.update(architecture='av_align', video_processing='features', audio_processing='features',
                  encoder_units_per_layer=(u3, u3), decoder_units_per_layer=(u3[0],), batch_size=(256, 256))
    else:
        raise ValueError(cfg)
    kw.update(over)
    return make_hparams(**kw)


def oracle_hparams(hp, model=None):
    """model: the product model whose generator words / stream ids the oracle should use (dropout, sampling).
    (The oracle is imported here, not at module level: bench.py's product arm uses this module's workload helpers and
    must not load the checker.)"""
    from oracle.avsr_oracle import OracleHParams
    rev = {v: k for k, v in hp.unit_dict.items()}
    rand = {}
    if model is not None:
        if hp.use_dropout:
            rand['dropout'] = dict(video=hp.video_encoder_dropout_probability,
                                   audio=hp.audio_encoder_dropout_probability,
                                   decoder=hp.decoder_dropout_probability)
        rand.update(sampling_probability_outputs=hp.sampling_probability_outputs, rng=model.rng_words(),
                    streams=model.random_streams)
    rand.update(regress_aus=bool(hp.regress_aus), au_loss_weight=hp.kwargs.get('au_loss_weight', 10.0),
                input_dense_layers=tuple(hp.input_dense_layers), residual_encoder=bool(hp.residual_encoder),
                highway_encoder=bool(hp.highway_encoder), enable_attention=hp.enable_attention is True,
                instance_normalisation=bool(hp.instance_normalisation), loss_fun=hp.loss_fun,
                encoder_weight_sharing=bool(hp.encoder_weight_sharing), label_smoothing=float(hp.label_smoothing),
                video_processing=hp.video_processing, cnn_filters=tuple(hp.kwargs.get('cnn_filters', (8, 16, 32, 64))))
    return OracleHParams(
        **rand,
        architecture=hp.architecture, encoder_type=hp.encoder_type,
        encoder_units_per_layer=hp.encoder_units_per_layer, decoder_units_per_layer=hp.decoder_units_per_layer,
        attention_type=hp.attention_type, embedding_size=hp.embedding_size, vocab_size=len(hp.unit_dict) - 1,
        go_id=rev['GO'], eos_id=rev['EOS'], batch_normalisation=hp.batch_normalisation,
        recurrent_l2_regularisation=hp.recurrent_l2_regularisation, clip_gradients=hp.clip_gradients,
        max_gradient_norm=hp.max_gradient_norm, learning_rate=hp.learning_rate,
        warmup_steps=hp.kwargs.get('warmup_steps', 750), beam_width=hp.beam_width,
        max_label_length=hp.max_label_length)


def synthetic_batch(hp, B, Ta=300, Tv=75, Fa=80, Fv=128, L=40, ragged=False, seed=0):
    """audio N(0,1) seed 1001; video U(-1,1) seed 1002; labels uniform 1..28 seed 1003; EOS=29 appended
    (io_utils.py:81-83).  ragged=True draws lengths in [T/2, T] and [L/2, L] (seed 1004)."""
    ra, rv, rl, rr = (np.random.default_rng(s + seed) for s in (1001, 1002, 1003, 1004))
    need_v = hp.video_processing is not None
    need_a = hp.audio_processing is not None
    lab_len = rr.integers(max(1, L // 2), L + 1, B) if ragged else np.full(B, L)
    labels = np.zeros((B, int(lab_len.max()) + 1), np.int32)
    for b in range(B):
        labels[b, :lab_len[b]] = rl.integers(1, 29, lab_len[b])
        labels[b, lab_len[b]] = 29
    lab_len = (lab_len + 1).astype(np.int32)
    out = {'labels': labels, 'labels_len': lab_len}
    if need_a:
        alen = rr.integers(max(1, Ta // 2), Ta + 1, B) if ragged else np.full(B, Ta)
        if ragged:
            alen[0] = Ta
        a = ra.standard_normal((B, Ta, Fa)).astype(np.float32)
        a *= (np.arange(Ta)[None, :, None] < alen[:, None, None])  # padded_batch pads with zeros
        out.update(audio=a, audio_len=alen.astype(np.int32))
    if need_v:
        vlen = rr.integers(max(1, Tv // 2), Tv + 1, B) if ragged else np.full(B, Tv)
        if ragged:
            vlen[0] = Tv
        v = rv.uniform(-1, 1, (B, Tv, Fv)).astype(np.float32)
        v *= (np.arange(Tv)[None, :, None] < vlen[:, None, None])
        out.update(video=v, video_len=vlen.astype(np.int32))
    return out


def to_image_sequences(batch, hw, channels=3, seed=1006):
    """Replaces the video features by lip crops [B,Tv,hw,hw,channels] ~ U(-1,1), zero past the length (resnet_cnn)."""
    B, Tv = batch['video'].shape[:2]
    v = np.random.default_rng(seed).uniform(-1, 1, (B, Tv, hw, hw, channels)).astype(np.float32)
    v *= (np.arange(Tv)[None, :, None, None, None] < batch['video_len'][:, None, None, None, None])
    batch['video'] = v
    return batch


def add_aus(batch, seed=1005):
    """Action-Unit payload of the video stream: [B,Tv,2] intensities in [-0.5, 4] (the loader clips to [0, 3])."""
    B, Tv = batch['video'].shape[:2]
    batch['aus'] = np.random.default_rng(seed).uniform(-0.5, 4.0, (B, Tv, 2)).astype(np.float32)
    return batch


def to_data_sequences(batch):
    def one(key):
        if key not in batch:
            return None
        payload = {'aus': batch['aus']} if key == 'video' and 'aus' in batch else None
        return make_batched_data(batch[key], batch[key + '_len'], batch['labels'], batch['labels_len'],
                                 payload=payload)
    return (one('video'), one('audio'))


def cast_batch(batch, dtype):
    return {k: (v.astype(dtype) if v.dtype.kind == 'f' else v) for k, v in batch.items()}
