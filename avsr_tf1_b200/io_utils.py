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


# --------------------------------------------------------------------------------------------------
# TFRecord iterators (io_utils.py:88-305 of the reference), without TensorFlow.
# The reference returns a BatchedData of graph nodes that session.run advances; here the same attribute names live
# on an eager iterator object: `next()` advances to the next batch (raising OutOfRangeError at the end of the epoch,
# like tf.errors.OutOfRangeError in avsr.py:300,492), `iterator_initializer()` restarts the epoch, and the
# BatchedData attributes (`inputs`, `inputs_length`, `labels`, ...) always describe the current batch.
# --------------------------------------------------------------------------------------------------
import queue
import threading

import numpy as np

SHUFFLE_BUFFER = 5000  # io_utils.py:102


class OutOfRangeError(Exception):
    """End of the epoch (tf.errors.OutOfRangeError)."""


class _HostRing(object):
    """Page-locked staging memory for the batches in flight: `depth` flat buffers per role (a role = one field of a
    batch, e.g. the features of stream 0), handed out round-robin as views of the requested shape.  A buffer is as
    large as the largest batch of its role seen so far (grown geometrically), so an epoch of bucketed batches with
    hundreds of distinct padded lengths pins depth x roles buffers, not one set per shape (cudaHostAlloc of a 300 MB
    batch costs more than decoding it, and pinned memory is never swapped).  A view is handed out again `depth` batches
    later: the consumer must have finished its host-to-device copy by then (the training loop reads the loss back every
    step, which synchronises)."""

    def __init__(self, pin, depth):
        self._pin, self._depth, self._rings = pin, max(2, int(depth)), {}

    def get(self, role, shape, dtype):
        import torch
        nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        ring = self._rings.setdefault(role, [[None] * self._depth, 0])
        i = ring[1]                       # slots are used in order 0, 1, ..., depth-1, 0, ...: reuse distance = depth
        ring[1] = (i + 1) % self._depth
        buf = ring[0][i]
        if buf is None or buf.numel() < nbytes:
            cap = nbytes if buf is None else max(nbytes, buf.numel() + buf.numel() // 2)
            buf = torch.empty(max(cap, 16), dtype=torch.uint8)
            if self._pin and torch.cuda.is_available():
                buf = buf.pin_memory()
            ring[0][i] = buf              # (a view of the replaced buffer that is still in use keeps it alive)
        return buf[:nbytes].view(dtype).view(tuple(shape))

    def pinned_bytes(self):
        return sum(b.numel() for ring in self._rings.values() for b in ring[0] if b is not None)


class _Shard(np.ndarray):
    """This rank's slice of a global batch (an index array) that remembers the global batch: every rank pads its input
    streams to the GLOBAL batch's longest sequence and reports the global batch size, so the all-reduced batch-norm
    sums of the ranks are sums over identically shaped slices of one batch (N ranks == one rank with the whole batch,
    also when the batch does not divide evenly)."""

    def __new__(cls, idx, global_idx):
        obj = np.asarray(idx, np.int64).view(cls)
        obj.global_idx = np.asarray(global_idx, np.int64)
        return obj

    def __array_finalize__(self, obj):
        self.global_idx = getattr(obj, 'global_idx', None)


class RecordBatcher(object):
    """zip(input record(s), label record) -> [filter] -> [shuffle(5000)] -> padded_batch, optionally bucketed with
    group_by_window(key = input_length // bucket_width, window_size = batch_size) -> prefetch.

    input_records: 0, 1 or 2 TFRecord paths (two: video first, then audio - io_utils.py:168); the bucketing key is
    the length of the FIRST stream (io_utils.py:125-129, 213-216).  Batches are assembled by the native library
    straight into pinned host tensors ([B,T,...] batch-major like the reference)."""

    def __init__(self, input_records, label_record, unit_dict, batch_size, shuffle=False, reverse_input=False,
                 bucket_width=-1, num_cores=4, max_sentence_length=None, seed=0, pin_memory=True, prefetch=2,
                 shuffle_buffer=SHUFFLE_BUFFER, shard=None):
        from .tfrecord import KIND_LABELS, RecordFile
        self._inputs = [RecordFile(p) for p in input_records]
        self._labels = RecordFile(label_record)
        if self._labels.kind != KIND_LABELS:
            raise Exception('%s is not a label record' % label_record)
        n = len(self._labels)
        for f in self._inputs:
            if f.kind == KIND_LABELS:
                raise Exception('%s is a label record' % f.path)
            n = min(n, len(f))  # Dataset.zip stops at the shortest component
        self._n = n
        ivdict = {v: k for k, v in unit_dict.items()}
        self._eos = ivdict['EOS']
        self.batch_size, self.shuffle, self.reverse_input = int(batch_size), bool(shuffle), bool(reverse_input)
        self.bucket_width, self.num_cores = int(bucket_width), int(num_cores)
        self.max_sentence_length = max_sentence_length
        self._rng = np.random.default_rng(seed)
        self._pin, self._prefetch, self._shuffle_buffer = pin_memory, int(prefetch), int(shuffle_buffer)
        # data parallelism (no counterpart in the single-device reference): every rank walks the SAME global batches
        # (same seed) and assembles only its contiguous slice of each; batches smaller than the world are dropped so
        # that all ranks take the same number of steps (their collectives must pair up)
        self._shard = None if shard is None or int(shard[1]) <= 1 else (int(shard[0]), int(shard[1]))
        self.has_aus = any(f.has_aus for f in self._inputs[:1])
        self._queue = self._thread = None
        self._current = None
        self._stop = threading.Event()
        self._ring = _HostRing(pin_memory, self._prefetch + 3)  # queue depth + one in assembly + one being consumed

    # ---- epoch order -----------------------------------------------------------------------------------
    def _element_order(self):
        idx = np.arange(self._n)
        if self.max_sentence_length is not None:  # io_utils.py:99-100: labels_length (with EOS) < max
            idx = idx[(self._labels.lengths[idx] + 1) < self.max_sentence_length]
        if not self.shuffle:
            return idx
        # tf.data shuffle: a buffer of 5000 elements, each output drawn uniformly from the buffer
        out, buf = np.empty(len(idx), np.int64), []
        k = 0
        for i in idx:
            buf.append(i)
            if len(buf) > self._shuffle_buffer:
                j = int(self._rng.integers(len(buf)))
                buf[j], buf[-1] = buf[-1], buf[j]
                out[k] = buf.pop()
                k += 1
        while buf:
            j = int(self._rng.integers(len(buf)))
            buf[j], buf[-1] = buf[-1], buf[j]
            out[k] = buf.pop()
            k += 1
        return out

    def batches_of_epoch(self):
        """List of index arrays, one per batch, in emission order (this rank's slice under data parallelism)."""
        batches = self._global_batches()
        if self._shard is None:
            return batches
        from .parallel import shard_batch
        r, w = self._shard
        out = []
        for idx in batches:
            if len(idx) < w:
                continue
            lo, hi = shard_batch(len(idx), r, w)
            out.append(_Shard(idx[lo:hi], idx))
        return out

    def _global_batches(self):
        order = self._element_order()
        B = self.batch_size
        if self.bucket_width == -1:
            return [order[i:i + B] for i in range(0, len(order), B)]  # drop_remainder=False
        key_len = self._inputs[0].lengths if self._inputs else self._labels.lengths + 1
        windows, out = {}, []
        for i in order:  # group_by_window: a window is emitted as soon as it holds batch_size elements
            w = windows.setdefault(int(key_len[i] // self.bucket_width), [])
            w.append(i)
            if len(w) == B:
                out.append(np.asarray(w, np.int64))
                w.clear()
        out.extend(np.asarray(w, np.int64) for w in windows.values() if w)  # leftovers at the end of the input
        return out

    # ---- batch assembly ----------------------------------------------------------------------------------
    def _assemble(self, idx):
        import torch
        n = len(idx)
        gidx = getattr(idx, 'global_idx', None)
        gidx = idx if gidx is None else gidx
        idx = np.asarray(idx, np.int64)
        streams = []
        for k, f in enumerate(self._inputs):
            t_pad = int(f.lengths[gidx].max())
            x = self._ring.get(('x', k), (n, t_pad, f.feat), torch.float32)
            lens = torch.empty(n, dtype=torch.int32)
            aus = self._ring.get(('aus', k), (n, t_pad, 2), torch.float32) if (k == 0 and f.has_aus) else None
            f.fill_inputs(idx, t_pad, x, lens, aus_dst=aus, reverse=self.reverse_input and len(self._inputs) == 1,
                          n_threads=self.num_cores)
            if len(f.input_shape) == 3:
                x = x.view(n, t_pad, *f.input_shape)
            names = np.array([f.filename(i) for i in idx], dtype=object)
            streams.append((x, lens, names, aus))
        l_pad = int(self._labels.lengths[idx].max()) + 1
        labels = torch.empty((n, l_pad), dtype=torch.int32)
        lab_len = torch.empty(n, dtype=torch.int32)
        self._labels.fill_labels(idx, l_pad, self._eos, labels, lab_len)
        lab_names = np.array([self._labels.filename(i) for i in idx], dtype=object)
        return streams, labels, lab_len, lab_names, int(len(gidx))

    def _producer(self, batches, q, stop):
        try:
            for idx in batches:
                if stop.is_set():
                    return
                q.put(self._assemble(idx))
            q.put(None)
        except BaseException as e:  # surfaced by next()
            q.put(e)

    def iterator_initializer(self):
        """Restart the epoch (session.run(iterator_initializer), avsr.py:262)."""
        self._shutdown()
        batches = self.batches_of_epoch()
        self._stop = threading.Event()
        if self._prefetch > 0:
            self._queue = queue.Queue(maxsize=self._prefetch)
            self._thread = threading.Thread(target=self._producer, args=(batches, self._queue, self._stop), daemon=True)
            self._thread.start()
        else:
            self._pending = iter(batches)
        self._current = None
        self._exhausted = False

    def _shutdown(self):
        if self._thread is not None:
            self._stop.set()
            while self._thread.is_alive():
                try:
                    self._queue.get(timeout=0.05)
                except queue.Empty:
                    pass
            self._thread = None
        self._queue = None

    def next(self):
        if getattr(self, '_exhausted', False):
            raise OutOfRangeError()  # stays at the end until the initializer runs again
        if self._queue is None and getattr(self, '_pending', None) is None:
            raise Exception('iterator is not initialised: call iterator_initializer() first')
        if self._queue is not None:
            item = self._queue.get()
            if isinstance(item, BaseException):
                self._current, self._exhausted = None, True  # the producer thread is gone: never block on its queue again
                raise item
        else:
            idx = next(self._pending, None)
            item = None if idx is None else self._assemble(idx)
        if item is None:
            self._current = None
            self._exhausted = True
            raise OutOfRangeError()
        self._current = item
        return self

    def __iter__(self):
        self.iterator_initializer()
        while True:
            try:
                yield self.next()
            except OutOfRangeError:
                return

    def __len__(self):
        return self._n

    # ---- BatchedData view of the current batch -------------------------------------------------------------
    def _cur(self):
        if self._current is None:
            raise OutOfRangeError()
        return self._current

    def _stream_field(self, j):
        streams = self._cur()[0]
        if not streams:
            return None
        vals = tuple(s[j] for s in streams)
        return vals[0] if len(vals) == 1 else vals

    inputs = property(lambda self: self._stream_field(0))
    inputs_length = property(lambda self: self._stream_field(1))
    inputs_filenames = property(lambda self: self._stream_field(2))
    labels = property(lambda self: self._cur()[1])
    labels_length = property(lambda self: self._cur()[2])
    labels_filenames = property(lambda self: self._cur()[3])

    @property
    def payload(self):
        streams = self._cur()[0]
        if streams and streams[0][3] is not None:
            return {'aus': streams[0][3]}
        return {}

    def data_sequences(self):
        """(video BatchedData | None, audio BatchedData | None) of the current batch - what avsr.py:566-570 hands to
        Seq2SeqModel (_parse_iterator / _parse_multimodal_iterator, avsr.py:573-626)."""
        from .tfrecord import KIND_VIDEO
        streams, labels, lab_len, lab_names, global_b = self._cur()
        # under data parallelism (no counterpart in the reference) the shard also names the size of its global batch
        extra = {'global_batch_size': global_b} if self._shard is not None else {}
        out = [None, None]
        for f, (x, lens, names, aus) in zip(self._inputs, streams):
            slot = 0 if (f.kind == KIND_VIDEO or (len(streams) == 2 and f is self._inputs[0])) else 1
            out[slot] = BatchedData(iterator_initializer=self.iterator_initializer, inputs=x, inputs_length=lens,
                                    inputs_filenames=names, labels=labels, labels_length=lab_len,
                                    labels_filenames=lab_names,
                                    payload=dict({'aus': aus} if aus is not None else {}, **extra))
        return tuple(out)


def make_iterator_from_one_record(data_record, label_record, unit_dict, batch_size, shuffle=False,
                                  reverse_input=False, bucket_width=-1, num_cores=4, max_sentence_length=None,
                                  **kw):
    """io_utils.py:88-165."""
    return RecordBatcher([data_record], label_record, unit_dict, batch_size, shuffle=shuffle,
                         reverse_input=reverse_input, bucket_width=bucket_width, num_cores=num_cores,
                         max_sentence_length=max_sentence_length, **kw)


def make_iterator_from_two_records(video_record, audio_record, label_record, batch_size, unit_dict, shuffle=False,
                                   reverse_input=False, bucket_width=-1, num_cores=4, **kw):
    """io_utils.py:168-257 (reverse_input is a TODO there and ignored; bucketing by the video length)."""
    return RecordBatcher([video_record, audio_record], label_record, unit_dict, batch_size, shuffle=shuffle,
                         reverse_input=False, bucket_width=bucket_width, num_cores=num_cores, **kw)


def make_iterator_from_label_record(label_record, batch_size, unit_dict, shuffle=False, reverse_input=False,
                                    bucket_width=-1, num_cores=4, **kw):
    """io_utils.py:260-305 (shuffle buffer 45000; bucketing by the label length incl. EOS)."""
    kw.setdefault('shuffle_buffer', 45000)
    return RecordBatcher([], label_record, unit_dict, batch_size, shuffle=shuffle, bucket_width=bucket_width,
                         num_cores=num_cores, **kw)


def _get_input_shape_from_record(record):
    """io_utils.py:308-341: ([input_size] | [width, height, channels], {'stream': 'feature' | 'video'[, 'aus': True]})."""
    from .tfrecord import KIND_FEATURE, KIND_LABELS, RecordFile
    f = RecordFile(record)
    try:
        if f.kind == KIND_LABELS:
            raise Exception('%s is a label record' % record)
        content_type = {'stream': 'feature' if f.kind == KIND_FEATURE else 'video'}
        if f.has_aus:
            content_type['aus'] = True
        return list(f.input_shape), content_type
    finally:
        f.close()


def _get_unit_from_record(record):
    """io_utils.py:344-351: the `unit` string stored in the context of a label record."""
    from .tfrecord import RecordFile
    f = RecordFile(record)
    try:
        return f.unit
    finally:
        f.close()
