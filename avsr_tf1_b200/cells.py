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
    cell_list = []
    for layer, units in enumerate(num_units_per_layer):
        if layer > 1 and weight_sharing is True:
            # cells.py:77-78: the SAME cell object is used again: layers 2.. run with the variables of layer 1 (the cell
            # is already built when MultiRNNCell calls it under `cell_2`); every position still draws its own masks
            cell = LSTMCellSpec(units, cell_list[-1].use_dropout, cell_list[-1].dropout_probability)
            cell.share_with = 1
            cell.residual = cell_list[-1].residual
            cell.highway = cell_list[-1].highway  # (the wrapper object is shared, its carry variables are per position)
        else:
            cell = _build_single_cell(cell_type, units, use_dropout, mode, dropout_probability, dtype)
            cell.share_with = None
            # cells.py:89-92: HighwayWrapper(cell) for layer > 0 (carry = sigmoid(x Wc + bc), output = x * carry + cell
            # output * (1 - carry)), else ResidualWrapper(cell): output = cell output + (un-dropped) layer input
            cell.highway = bool(highway_connections is True and layer > 0)
            cell.residual = bool(residual_connections is True and layer > 0 and not cell.highway)
        cell_list.append(cell)
    if len(cell_list) == 1:
        return cell_list[0]
    return cell_list
