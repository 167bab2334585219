"""Attention mechanism factory - drop-in for reference avsr/attention.py.

Supported scorers: bahdanau, normed_bahdanau, luong, scaled_luong
(attention.py:25-72).  The monotonic variants (attention.py:43-54, 73-84) and
`deep_fusion` (attention.py:120-122) are never selected by any reference script
and raise here."""
from __future__ import annotations

from .layers import AttnLSTMOp, BuildContext, MechDef

_SUPPORTED = ('bahdanau', 'normed_bahdanau', 'luong', 'scaled_luong')


def create_attention_mechanism(attention_type, num_units, memory_depth, ctx: BuildContext, wrap_prefix, idx,
                               mem_layer_name, query_depth):
    """Returns (MechDef, output_attention) like attention.py:5-88."""
    if attention_type in ('normed_monotonic_bahdanau', 'scaled_monotonic_luong'):
        raise NotImplementedError('monotonic attention is not implemented on the B200 path')
    if attention_type not in _SUPPORTED:
        raise Exception('unknown attention mechanism')
    if 'luong' in attention_type and num_units != query_depth:
        raise ValueError('Luong attention requires num_units (%d) == query depth (%d)' % (num_units, query_depth))
    md = MechDef(ctx, attention_type, wrap_prefix, idx, mem_layer_name, query_depth, memory_depth, num_units)
    return md, md.output_attention


def create_attention_mechanisms(num_units, attention_types, memory_depths, ctx, wrap_prefix, mem_layer_names,
                                query_depth, fusion_type='linear_fusion'):
    """attention.py:91-129 (linear_fusion: one Dense(num_units) attention layer per mechanism)."""
    if fusion_type == 'deep_fusion':
        raise NotImplementedError('deep_fusion is never selected by the reference callers')
    if fusion_type != 'linear_fusion':
        raise Exception('Unknown fusion type')
    mechanisms, output_attention = [], None
    for idx, (attention_type, depth) in enumerate(zip(attention_types, memory_depths)):
        md, output_attention = create_attention_mechanism(attention_type, num_units, depth, ctx, wrap_prefix, idx,
                                                          mem_layer_names[idx], query_depth)
        mechanisms.append(md)
    return mechanisms, output_attention


def add_attention(cell, attention_types, num_units, memory_depths, ctx, wrap_prefix, mem_layer_names, in_dim,
                  fusion_type='linear_fusion'):
    """Wraps `cell` (an LSTMCellSpec) into an AttentionWrapper (attention.py:132-191).
    Returns the AttnLSTMOp that runs the wrapped cell over a whole sequence."""
    mechs, _ = create_attention_mechanisms(num_units, attention_types, memory_depths, ctx, wrap_prefix,
                                           mem_layer_names, cell.num_units, fusion_type)
    return AttnLSTMOp(ctx, wrap_prefix, in_dim, cell.num_units, mechs, drop=ctx.drop_state(cell, wrap_prefix))
