"""ctypes binding of include/avsr_io.h (libavsr_io.so): TFRecord / SequenceExample files as the reference's
dataset_writer.py writes them, read without TensorFlow.  Host-only; see io_utils.py for the batching iterators."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('AVSR_IO_LIB') or os.path.join(_HERE, 'lib', 'libavsr_io.so')

KIND_FEATURE, KIND_VIDEO, KIND_LABELS = 0, 1, 2


class AvsrIoInfo(C.Structure):
    _fields_ = [('kind', C.c_int), ('has_aus', C.c_int), ('n_records', C.c_longlong), ('feat', C.c_longlong),
                ('width', C.c_int), ('height', C.c_int), ('channels', C.c_int), ('unit', C.c_char * 32)]


_P, _I, _L = C.c_void_p, C.c_int, C.c_longlong
PROTOTYPES = {
    'avsr_io_last_error': (C.c_char_p, []),
    'avsr_io_crc32c': (C.c_uint32, [_P, C.c_size_t]),
    'avsr_io_masked_crc32c': (C.c_uint32, [_P, C.c_size_t]),
    'avsr_io_open': (_I, [C.c_char_p, _I, C.POINTER(_P)]),
    'avsr_io_close': (None, [_P]),
    'avsr_io_info': (_I, [_P, C.POINTER(AvsrIoInfo)]),
    'avsr_io_lengths': (_I, [_P, _P]),
    'avsr_io_filename': (_I, [_P, _L, C.c_char_p, _I]),
    'avsr_io_fill_inputs': (_I, [_P, _P, _I, _I, _P, _P, _P, _I, _I]),
    'avsr_io_fill_labels': (_I, [_P, _P, _I, _I, C.c_int32, _P, _P]),
    'avsr_io_writer_open': (_I, [C.c_char_p, C.POINTER(_P)]),
    'avsr_io_writer_close': (_I, [_P]),
    'avsr_io_write_feature': (_I, [_P, C.c_char_p, _P, _I, _I]),
    'avsr_io_write_video': (_I, [_P, C.c_char_p, _P, _I, _I, _I, _I, _P]),
    'avsr_io_write_labels': (_I, [_P, C.c_char_p, _P, _I, C.c_char_p]),
}

_lib = None


class AvsrIoError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AvsrIoError(f'{LIB_PATH} not found: build it first (make)')
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise AvsrIoError(load().avsr_io_last_error().decode('utf-8', 'replace'))


def crc32c(data: bytes) -> int:
    return int(load().avsr_io_crc32c(data, len(data)))


def masked_crc32c(data: bytes) -> int:
    return int(load().avsr_io_masked_crc32c(data, len(data)))


def _ptr(a):
    """numpy array or torch tensor (host) -> address."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()


class RecordFile(object):
    """One TFRecord file of SequenceExamples, indexed at open (include/avsr_io.h avsr_io_open)."""

    def __init__(self, path, verify_data=False):
        self.path = str(path)
        h = _P()
        check(load().avsr_io_open(self.path.encode(), int(bool(verify_data)), C.byref(h)))
        self._h = h
        info = AvsrIoInfo()
        check(load().avsr_io_info(self._h, C.byref(info)))
        self.kind, self.has_aus, self.n = info.kind, bool(info.has_aus), int(info.n_records)
        self.feat = int(info.feat)
        self.unit = info.unit.decode()
        # _get_input_shape_from_record (io_utils.py:308-341): [input_size] or [width, height, channels]
        self.input_shape = [self.feat] if self.kind == KIND_FEATURE else [info.width, info.height, info.channels]
        self.lengths = np.zeros(self.n, np.int64)
        if self.n:
            check(load().avsr_io_lengths(self._h, self.lengths.ctypes.data))

    def __len__(self):
        return self.n

    def close(self):
        if self._h:
            load().avsr_io_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def filename(self, idx) -> bytes:
        buf = C.create_string_buffer(512)
        check(load().avsr_io_filename(self._h, int(idx), buf, 512))
        return buf.value

    def fill_inputs(self, idx, t_pad, dst, lens, aus_dst=None, reverse=False, n_threads=4):
        """dst [n, t_pad, feat] float32, lens [n] int32 (numpy or host torch tensors, C-contiguous)."""
        idx = np.ascontiguousarray(idx, np.int64)
        check(load().avsr_io_fill_inputs(self._h, idx.ctypes.data, len(idx), int(t_pad), _ptr(dst), _ptr(aus_dst),
                                         _ptr(lens), int(bool(reverse)), int(n_threads)))

    def fill_labels(self, idx, l_pad, eos, dst, lens):
        idx = np.ascontiguousarray(idx, np.int64)
        check(load().avsr_io_fill_labels(self._h, idx.ctypes.data, len(idx), int(l_pad), int(eos), _ptr(dst),
                                         _ptr(lens)))


class RecordWriter(object):
    """Writes the reference's three example kinds (dataset_writer.py:290-311, 439-458, 461-498)."""

    def __init__(self, path):
        h = _P()
        check(load().avsr_io_writer_open(str(path).encode(), C.byref(h)))
        self._h = h

    def write_feature(self, sentence_id: str, inputs):
        x = np.ascontiguousarray(inputs, np.float32)
        check(load().avsr_io_write_feature(self._h, sentence_id.encode(), x.ctypes.data, x.shape[0], x.shape[1]))

    def write_video(self, sentence_id: str, frames, aus=None):
        x = np.ascontiguousarray(frames, np.float32)
        if x.ndim == 3:
            x = x[..., None]
        T, Hh, Ww, Cc = x.shape
        a = None if aus is None else np.ascontiguousarray(aus, np.float32)
        check(load().avsr_io_write_video(self._h, sentence_id.encode(), x.ctypes.data, T, Hh, Ww, Cc,
                                         None if a is None else a.ctypes.data))

    def write_labels(self, label_id: str, labels, unit='character'):
        y = np.ascontiguousarray(labels, np.int64)
        check(load().avsr_io_write_labels(self._h, label_id.encode(), y.ctypes.data, len(y), unit.encode()))

    def close(self):
        if self._h:
            h, self._h = self._h, None
            check(load().avsr_io_writer_close(h))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
