"""ctypes binding of the C ABI in include/avsr_b200.h (libavsr_b200.so).

There is no CPU fallback: if the shared library is missing, importing any op
raises.  Build it with ``make`` (or ``python -c 'import __graft_entry__ as g; g.build()'``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('AVSR_B200_LIB') or os.path.join(_HERE, 'lib', 'libavsr_b200.so')

c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int)


class AvsrAttnMech(C.Structure):
    _fields_ = [
        ('kind', C.c_int), ('Tm', C.c_int), ('Dm', C.c_int), ('A', C.c_int),
        ('values', C.c_void_p), ('keys', C.c_void_p), ('mem_len', C.c_void_p),
        ('Wl', C.c_void_p), ('Wq', C.c_void_p), ('v', C.c_void_p), ('g', C.c_void_p), ('bias', C.c_void_p),
        ('align', C.c_void_p), ('hc', C.c_void_p), ('pq', C.c_void_p),
        ('dkeys', C.c_void_p), ('dvalues', C.c_void_p), ('dWl', C.c_void_p), ('dWq', C.c_void_p),
        ('dv', C.c_void_p), ('dg', C.c_void_p), ('dbias', C.c_void_p), ('dpq', C.c_void_p),
        ('ds', C.c_void_p), ('dhc', C.c_void_p), ('values_op', C.c_void_p),
    ]


class AvsrSampling(C.Structure):
    _fields_ = [
        ('Wd', C.c_void_p), ('bd', C.c_void_p), ('embedding', C.c_void_p), ('Wx', C.c_void_p), ('bias', C.c_void_p),
        ('used_ids', C.c_void_p), ('sample_ids', C.c_void_p), ('x', C.c_void_p),
        ('V', C.c_int), ('E', C.c_int), ('stream', C.c_uint32), ('thr_p', C.c_uint32),
    ]


class AvsrRnnSeq(C.Structure):
    _fields_ = [
        ('T', C.c_int), ('B', C.c_int), ('H', C.c_int), ('n_mech', C.c_int), ('output_attention', C.c_int),
        ('len', C.c_void_p), ('gates', C.c_void_p), ('Wrec', C.c_void_p), ('c0', C.c_void_p),
        ('S', C.c_void_p), ('craw', C.c_void_p), ('out', C.c_void_p), ('cT', C.c_void_p), ('hT', C.c_void_p),
        ('mech', AvsrAttnMech * 2),
        ('dout', C.c_void_p), ('dcT', C.c_void_p), ('dhT', C.c_void_p), ('dZ', C.c_void_p), ('dA', C.c_void_p),
        ('dWrec', C.c_void_p), ('dc0', C.c_void_p), ('dh0', C.c_void_p), ('dbias', C.c_void_p), ('work', C.c_void_p),
        ('grad_scale', C.c_float),
        ('rng', C.c_void_p), ('drop_stream', C.c_uint32), ('thr_in', C.c_uint32), ('thr_state', C.c_uint32),
        ('thr_out', C.c_uint32), ('t_begin', C.c_int), ('t_end', C.c_int), ('stepwise', C.c_int),
        ('samp', C.POINTER(AvsrSampling)),
    ]


ATTN_KINDS = {'luong': 0, 'scaled_luong': 1, 'bahdanau': 2, 'normed_bahdanau': 3}

_P, _I, _L, _F, _D, _U = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_double, C.c_uint32

# name -> (restype, argtypes).  Must list every symbol include/avsr_b200.h declares
# (tests/test_abi.py checks both directions).
PROTOTYPES = {
    'avsr_last_error': (C.c_char_p, []),
    'avsr_version': (_I, []),
    'avsr_launch_count': (C.c_ulonglong, []),
    'avsr_kernel_timing': (_I, [_I]),
    'avsr_kernel_times': (_I, [_P, _P]),
    'avsr_set_tensor_cores': (_I, [_I]),
    'avsr_get_tensor_cores': (_I, []),
    'avsr_round_tf32': (_I, [_P, _P, _P, _L]),
    'avsr_gemm': (_I, [_P, _I, _I, _I, _I, _I, _P, _I, _P, _I, _P, _I, _F, _P, _I]),
    'avsr_colsum': (_I, [_P, _P, _I, _I, _I, _P]),
    'avsr_bn_stats': (_I, [_P, _P, _L, _I, _P]),
    'avsr_bn_apply_train': (_I, [_P, _P, _L, _I, _P, _D, _P, _P, _F, _F, _P, _P, _P, _P, _P]),
    'avsr_bn_apply_train_t': (_I, [_P, _P, _I, _I, _I, _P, _D, _P, _P, _F, _F, _P, _P, _P, _P, _P]),
    'avsr_bn_input_grads': (_I, [_P, _P, _I, _P, _I, _P, _P, _P, _I, _I, _P, _P]),
    'avsr_bn_apply_eval': (_I, [_P, _P, _L, _I, _P, _P, _P, _P, _F, _P]),
    'avsr_bn_bwd_stats': (_I, [_P, _P, _P, _L, _I, _P]),
    'avsr_bn_bwd_apply': (_I, [_P, _P, _P, _L, _I, _P, _D, _P, _P, _P, _P, _P]),
    'avsr_reverse_sequence': (_I, [_P, _P, _P, _I, _I, _I, _P]),
    'avsr_transpose01': (_I, [_P, _P, _P, _I, _I, _I]),
    'avsr_u8_to_f32': (_I, [_P, _P, _L, _F, _F, _P]),
    'avsr_rnn_work_floats': (C.c_size_t, [_I, _I, _I, _I, _I, _I]),
    'avsr_struct_sizes': (_I, [C.POINTER(C.c_int)]),
    'avsr_rnn_sampling_fused': (_I, [C.POINTER(AvsrRnnSeq)]),
    'avsr_rnn_seq_fwd': (_I, [_P, C.POINTER(AvsrRnnSeq)]),
    'avsr_rnn_seq_bwd': (_I, [_P, C.POINTER(AvsrRnnSeq)]),
    'avsr_normed_v_fwd': (_I, [_P, _P, _P, _I, _P]),
    'avsr_normed_v_bwd': (_I, [_P, _P, _P, _P, _I, _P, _P]),
    'avsr_dropout': (_I, [_P, _P, _L, _L, _P, _U, _U, _I, _P]),
    'avsr_sched_sample': (_I, [_P, _P, _I, _I, _P, _U, _I, _U, _P, _P, _P]),
    'avsr_embedding_fwd': (_I, [_P, _P, _I, _I, _P, _L, _P]),
    'avsr_embedding_bwd': (_I, [_P, _P, _P, _L, _I, _I, _P]),
    'avsr_seq_loss': (_I, [_P, _P, _I, _I, _I, _P, _I, _P, _P, _F, _P, _P]),
    'avsr_seq_loss_devel': (_I, [_P, _P, _I, _I, _I, _P, _I, _P, _P, _I, _F, _P, _P]),
    'avsr_au_loss': (_I, [_P, _P, _I, _I, _P, _P, _P, _P, _P]),
    'avsr_sumsq': (_I, [_P, _P, _L, _P]),
    'avsr_axpy': (_I, [_P, _F, _P, _P, _L]),
    'avsr_adam_clip_step': (_I, [_P, _P, _P, _P, _P, _L, _P, _F, _P, _F, _F, _F, _P]),
    'avsr_optim_clip_step': (_I, [_P, _I, _P, _P, _P, _P, _L, _P, _F, _P, _F, _F, _F, _F, _P]),
    'avsr_im2col': (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    'avsr_col2im': (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    'avsr_conv2d_direct': (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    'avsr_conv2d_wgrad': (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    'avsr_conv2d_tc_supported': (_I, [_I, _I, _I, _I, _I]),
    'avsr_conv2d_tc': (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    'avsr_conv2d_wgrad_tc': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    'avsr_bn_finalize': (_I, [_P, _P, _D, _P, _P, _F, _F, _I, _P, _P, _P]),
    'avsr_bn_coef_eval': (_I, [_P, _P, _P, _P, _P, _F, _I, _P]),
    'avsr_bn_relu_bwd_apply': (_I, [_P, _P, _P, _P, _P, _D, _P, _L, _I, _P]),
    'avsr_relu_fwd': (_I, [_P, _P, _L, _P]),
    'avsr_relu_bwd': (_I, [_P, _P, _P, _L, _P]),
    'avsr_selu_fwd': (_I, [_P, _P, _L, _P]),
    'avsr_selu_bwd': (_I, [_P, _P, _P, _L, _P]),
    'avsr_highway_fwd': (_I, [_P, _P, _P, _P, _L, _P, _P]),
    'avsr_highway_bwd': (_I, [_P, _P, _P, _P, _P, _L, _P, _P, _P]),
    'avsr_greedy_pick': (_I, [_P, _P, _I, _I, _I, _P, _P, _P]),
    'avsr_beam_step': (_I, [_P, _P, _I, _I, _I, _I, _F, _P, _P, _P, _P, _P, _P]),
    'avsr_gather_rows': (_I, [_P, _P, _P, _L, _I, _P]),
}

_lib = None


class AvsrError(RuntimeError):
    pass


def load():
    """Load libavsr_b200.so (once).  Raises if it has not been built - by design."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AvsrError(f'{LIB_PATH} not found: build the CUDA library first (make). '
                        'There is deliberately no CPU fallback.')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    sizes = (C.c_int * 2)()
    lib.avsr_struct_sizes(sizes)
    if (sizes[0], sizes[1]) != (C.sizeof(AvsrAttnMech), C.sizeof(AvsrRnnSeq)):
        raise AvsrError('struct layout mismatch between include/avsr_b200.h and _lib.py: C %s vs ctypes %s'
                        % ((sizes[0], sizes[1]), (C.sizeof(AvsrAttnMech), C.sizeof(AvsrRnnSeq))))
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        raise AvsrError(load().avsr_last_error().decode('utf-8', 'replace'))
