"""avsr_tf1_b200 - B200-native implementation of the seq2seq hot path of
georgesterpu/avsr-tf1 (the reference package directory is `avsr/`; this package keeps
its module names: cells, encoder, attention, decoder_unimodal, decoder_bimodal, seq2seq).

The directory is named avsr_tf1_b200 (underscore) because `avsr-tf1_b200` is not an
importable Python identifier."""
from .hparams import HParams, create_unit_dict, make_hparams  # noqa: F401
from .io_utils import BatchedData, make_batched_data  # noqa: F401


def __getattr__(name):
    if name == 'Seq2SeqModel':  # lazy: importing the model loads libavsr_b200.so
        from .seq2seq import Seq2SeqModel
        return Seq2SeqModel
    raise AttributeError(name)
