"""BatchedData: the tuple the reference hands to Seq2SeqModel (io_utils.py:8-18).
Here the fields hold concrete batches (numpy / torch, host or device) instead of
tf.data iterator nodes."""
from __future__ import annotations

import collections

from .hparams import create_unit_dict  # noqa: F401  (re-export, io_utils.py:354)


class BatchedData(collections.namedtuple("BatchedData",
                                         ("iterator_initializer", "inputs", "inputs_length", "inputs_filenames",
                                          "labels", "labels_length", "labels_filenames", "payload"))):
    pass


def make_batched_data(inputs, inputs_length, labels, labels_length, filenames=None, payload=None):
    return BatchedData(iterator_initializer=None, inputs=inputs, inputs_length=inputs_length,
                       inputs_filenames=filenames, labels=labels, labels_length=labels_length,
                       labels_filenames=filenames, payload=payload or {})
