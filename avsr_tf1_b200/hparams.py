"""HParams stand-in for tf.contrib.training.HParams with the defaults of
AVSR.__init__ (reference avsr/avsr.py:21-75, frozen at :150-198)."""
from __future__ import annotations

import os

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_UNIT_FILE = os.path.join(_HERE, 'misc', 'character_list')


class HParams(object):
    def __init__(self, **kwargs):
        self.__dict__.update(kwargs)

    def values(self):
        return dict(self.__dict__)

    def override(self, **kwargs):
        d = dict(self.__dict__)
        d.update(kwargs)
        return HParams(**d)

    def __repr__(self):
        return 'HParams(%s)' % ', '.join('%s=%r' % kv for kv in sorted(self.__dict__.items()) if kv[0] != 'unit_dict')


def create_unit_dict(unit_file=None, unit_list=None):
    """io_utils.py:354-370.  Returns {id: symbol}: 0 MASK, -1 END, 1..n units, n+1 EOS, n+2 GO."""
    unit_dict = {'MASK': 0, 'END': -1}
    if unit_list is None:
        with open(unit_file or DEFAULT_UNIT_FILE, 'r') as f:
            unit_list = f.read().splitlines()
    idx = 0
    for idx, subunit in enumerate(unit_list):
        unit_dict[subunit] = idx + 1
    unit_dict['EOS'] = idx + 2
    unit_dict['GO'] = idx + 3
    return {v: k for k, v in unit_dict.items()}


def make_hparams(unit='character', unit_file=None, video_processing=None, audio_processing=None,
                 batch_size=(64, 64), regress_aus=False, batch_normalisation=True, instance_normalisation=False,
                 input_dense_layers=(0,), architecture='unimodal', encoder_type='unidirectional',
                 highway_encoder=False, residual_encoder=False, cell_type='lstm',
                 recurrent_l2_regularisation=0.0001, weight_decay=0.0001,
                 encoder_units_per_layer=((256,), (256, 256, 256)), decoder_units_per_layer=(256,),
                 encoder_weight_sharing=False, enable_attention=True,
                 attention_type=(('scaled_luong',) * 1, ('scaled_luong',) * 1), use_dropout=True,
                 audio_encoder_dropout_probability=(0.9, 0.9, 0.9),
                 video_encoder_dropout_probability=(0.9, 0.9, 0.9), decoder_dropout_probability=(0.9, 0.9, 0.9),
                 embedding_size=128, sampling_probability_outputs=0.1, label_smoothing=0.0,
                 decoding_algorithm='beam_search', beam_width=10, max_sentence_length=None, optimiser='Adam',
                 learning_rate=0.001, lr_decay=None, loss_fun=None, clip_gradients=True, max_gradient_norm=1.0,
                 num_gpus=1, write_attention_alignment=False, precision='float32', profiling=False, unit_dict=None,
                 **kwargs):
    """Same keyword surface and defaults as AVSR.__init__ (avsr.py:21-75)."""
    if unit_dict is None:
        unit_dict = create_unit_dict(unit_file)
    if precision != 'float32':
        raise ValueError('only float32 is supported on this path (the reference float16 path is broken, '
                         'seq2seq.py:234-240)')
    return HParams(
        unit_dict=unit_dict, unit_file=unit_file, vocab_size=len(unit_dict), batch_size=batch_size,
        video_processing=video_processing, audio_processing=audio_processing,
        max_label_length={'viseme': 150, 'phoneme': 150, 'character': 150}[unit],
        max_sentence_length=max_sentence_length, batch_normalisation=batch_normalisation,
        instance_normalisation=instance_normalisation, input_dense_layers=input_dense_layers,
        encoder_type=encoder_type, architecture=architecture, highway_encoder=highway_encoder,
        residual_encoder=residual_encoder, regress_aus=regress_aus, cell_type=cell_type,
        recurrent_l2_regularisation=None if optimiser == 'AdamW' else recurrent_l2_regularisation,
        weight_decay=weight_decay, encoder_units_per_layer=encoder_units_per_layer,
        decoder_units_per_layer=decoder_units_per_layer, encoder_weight_sharing=encoder_weight_sharing,
        bijective_state_copy=False, enable_attention=enable_attention, attention_type=attention_type,
        use_dropout=use_dropout, audio_encoder_dropout_probability=audio_encoder_dropout_probability,
        video_encoder_dropout_probability=video_encoder_dropout_probability,
        decoder_dropout_probability=decoder_dropout_probability, embedding_size=embedding_size,
        sampling_probability_outputs=sampling_probability_outputs, label_smoothing=label_smoothing,
        decoding_algorithm=decoding_algorithm, beam_width=beam_width, use_ctc=False, optimiser=optimiser,
        loss_scaling=1, learning_rate=learning_rate, lr_decay=lr_decay, loss_fun=loss_fun,
        clip_gradients=clip_gradients, max_gradient_norm=max_gradient_norm, num_gpus=num_gpus,
        write_attention_alignment=write_attention_alignment, dtype='float32', profiling=profiling, kwargs=kwargs)
