"""RNN cell factory - drop-in for reference avsr/cells.py.

Only the cell every shipped experiment uses is implemented on the B200 path:
``LSTMCell(use_peepholes=False, cell_clip=1.0)`` (cells.py:14-18).  The other
``cell_type`` strings of the reference (cells.py:19-42) raise the reference's own
exception text; they are out of scope."""
from __future__ import annotations


class LSTMCellSpec(object):
    """Describes one LSTMCell (+ optional DropoutWrapper, cells.py:46-54)."""

    def __init__(self, num_units, use_dropout=False, dropout_probability=(1.0, 1.0, 1.0)):
        self.num_units = int(num_units)
        self.use_dropout = bool(use_dropout)
        self.dropout_probability = tuple(dropout_probability)


def _build_single_cell(cell_type, num_units, use_dropout, mode, dropout_probability, dtype=None, device=None):
    if cell_type != 'lstm':
        raise Exception('cell type not supported: {}'.format(cell_type))
    # DropoutWrapper(input, state, output keep probabilities) in train mode only (cells.py:46-54).  TF's Philox
    # streams are not reproducible: the masks come from the library's counter-based generator (avsr_dropout), the
    # oracle restates it, parity holds mask for mask.
    drop = use_dropout is True and mode == 'train' and any(p < 1.0 for p in dropout_probability)
    return LSTMCellSpec(num_units, drop, dropout_probability)


def build_rnn_layers(cell_type, num_units_per_layer, use_dropout, dropout_probability, mode, dtype=None,
                     residual_connections=False, highway_connections=False, weight_sharing=False, as_list=False):
    """Same signature and return convention as cells.py:61-102: one cell for a single
    layer, else the stack (a list stands in for MultiRNNCell)."""
    if residual_connections or highway_connections or weight_sharing:
        raise NotImplementedError('residual / highway / weight-sharing encoders are off in every reference '
                                  'config (avsr.py:41-48) and not implemented on the B200 path')
    cell_list = [_build_single_cell(cell_type, units, use_dropout, mode, dropout_probability, dtype)
                 for units in num_units_per_layer]
    if len(cell_list) == 1:
        return cell_list[0]
    return cell_list
