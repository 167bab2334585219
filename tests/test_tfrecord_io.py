"""Input pipeline in front of the hot path: TFRecord framing + SequenceExample wire format of
the reference's dataset_writer.py, read and written by libavsr_io.so (include/avsr_io.h), and the eager mirrors of
io_utils.make_iterator_from_*.  Independent implementations pin the bytes: google.protobuf (messages declared from
tensorflow/core/example/{feature,example}.proto's published schema) and tensorboard's TFRecord reader / crc32c."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from avsr_tf1_b200 import io_utils, tfrecord
from avsr_tf1_b200.hparams import create_unit_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_symbol_of_the_header():
    hdr = open(os.path.join(ROOT, 'include', 'avsr_io.h')).read()
    declared = set(re.findall(r'\b(avsr_io_\w+)\s*\(', hdr))
    assert declared == set(tfrecord.PROTOTYPES), declared ^ set(tfrecord.PROTOTYPES)
    lib = C.CDLL(tfrecord.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name


def test_crc32c_known_answers():
    # RFC 3720 B.4 test vectors + the classic check value
    assert tfrecord.crc32c(b'123456789') == 0xE3069283
    assert tfrecord.crc32c(bytes(32)) == 0x8A9136AA
    assert tfrecord.crc32c(b'\xff' * 32) == 0x62A8AB43
    assert tfrecord.crc32c(bytes(range(32))) == 0x46DD794E
    assert tfrecord.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    assert tfrecord.crc32c(b'') == 0
    tb = pytest.importorskip('tensorboard.compat.tensorflow_stub.pywrap_tensorflow')
    rng = np.random.default_rng(0)
    for n in (1, 7, 8, 9, 63, 64, 65, 1000, 4097):
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert tfrecord.crc32c(data) == tb.crc32c(data)
        assert tfrecord.masked_crc32c(data) == tb.masked_crc32c(data)


def _example_classes():
    """tf.train.SequenceExample & co. declared with protobuf's descriptor API (no TensorFlow)."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    F = descriptor_pb2.FieldDescriptorProto
    fd = descriptor_pb2.FileDescriptorProto(name='avsr_test_example.proto', package='avsrtest', syntax='proto3')

    def msg(name):
        m = fd.message_type.add()
        m.name = name
        return m

    def field(m, name, num, typ, label=F.LABEL_OPTIONAL, type_name=None, packed=None, oneof=None):
        f = m.field.add(name=name, number=num, type=typ, label=label)
        if type_name:
            f.type_name = '.avsrtest.' + type_name
        if packed is not None:
            f.options.packed = packed
        if oneof is not None:
            f.oneof_index = oneof
        return f

    def map_field(m, name, num, value_type):
        e = m.nested_type.add(name=name.title().replace('_', '') + 'Entry')
        e.options.map_entry = True
        e.field.add(name='key', number=1, type=F.TYPE_STRING, label=F.LABEL_OPTIONAL)
        e.field.add(name='value', number=2, type=F.TYPE_MESSAGE, label=F.LABEL_OPTIONAL,
                    type_name='.avsrtest.' + value_type)
        m.field.add(name=name, number=num, type=F.TYPE_MESSAGE, label=F.LABEL_REPEATED,
                    type_name=f'.avsrtest.{m.name}.{e.name}')

    field(msg('BytesList'), 'value', 1, F.TYPE_BYTES, F.LABEL_REPEATED)
    field(msg('FloatList'), 'value', 1, F.TYPE_FLOAT, F.LABEL_REPEATED, packed=True)
    field(msg('Int64List'), 'value', 1, F.TYPE_INT64, F.LABEL_REPEATED, packed=True)
    feat = msg('Feature')
    feat.oneof_decl.add(name='kind')
    field(feat, 'bytes_list', 1, F.TYPE_MESSAGE, type_name='BytesList', oneof=0)
    field(feat, 'float_list', 2, F.TYPE_MESSAGE, type_name='FloatList', oneof=0)
    field(feat, 'int64_list', 3, F.TYPE_MESSAGE, type_name='Int64List', oneof=0)
    map_field(msg('Features'), 'feature', 1, 'Feature')
    field(msg('FeatureList'), 'feature', 1, F.TYPE_MESSAGE, F.LABEL_REPEATED, type_name='Feature')
    map_field(msg('FeatureLists'), 'feature_list', 1, 'FeatureList')
    se = msg('SequenceExample')
    field(se, 'context', 1, F.TYPE_MESSAGE, type_name='Features')
    field(se, 'feature_lists', 2, F.TYPE_MESSAGE, type_name='FeatureLists')
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName('avsrtest.SequenceExample'))


def _read_raw_records(path):
    tb = pytest.importorskip('tensorboard.compat.tensorflow_stub.pywrap_tensorflow')
    r = tb.PyRecordReader_New(path)
    out = []
    while True:
        try:
            r.GetNext()
        except Exception:
            break
        out.append(r.record())
    return out


def test_writer_bytes_parse_with_protobuf_and_tensorboard(tmp_path):
    SequenceExample = _example_classes()
    rng = np.random.default_rng(1)
    p = str(tmp_path / 'video.tfrecord')
    frames = rng.uniform(-1, 1, (5, 4, 3, 2)).astype(np.float32)  # [T, height, width, channels]
    aus = rng.uniform(0, 3, (5, 2)).astype(np.float32)
    with tfrecord.RecordWriter(p) as w:
        w.write_video('utt000001', frames, aus)
    with tfrecord.RecordWriter(str(tmp_path / 'labels.tfrecord')) as w:
        w.write_labels('utt000001', [3, 1, 4, 1, 5, 28], unit='character')
    recs = _read_raw_records(p)  # tensorboard verifies both masked CRCs of the framing
    assert len(recs) == 1
    ex = SequenceExample.FromString(recs[0])
    ctx = ex.context.feature
    assert ctx['input_length'].int64_list.value[0] == 5 and ctx['width'].int64_list.value[0] == 3
    assert ctx['height'].int64_list.value[0] == 4 and ctx['channels'].int64_list.value[0] == 2
    assert ctx['filename'].bytes_list.value[0] == b'utt000001'
    fl = ex.feature_lists.feature_list['inputs'].feature
    assert len(fl) == 5
    got = np.array([f.float_list.value for f in fl], np.float32)
    assert np.array_equal(got, frames.reshape(5, -1))
    assert np.array_equal(np.array([f.float_list.value for f in ex.feature_lists.feature_list['aus'].feature],
                                   np.float32), aus)
    lab = SequenceExample.FromString(_read_raw_records(str(tmp_path / 'labels.tfrecord'))[0])
    assert lab.context.feature['unit'].bytes_list.value[0] == b'character'
    assert lab.context.feature['labels_length'].int64_list.value[0] == 6
    assert [f.int64_list.value[0] for f in lab.feature_lists.feature_list['labels'].feature] == [3, 1, 4, 1, 5, 28]


def _frame(payload: bytes) -> bytes:
    tb = pytest.importorskip('tensorboard.compat.tensorflow_stub.pywrap_tensorflow')
    n = np.uint64(len(payload)).tobytes()
    return (n + np.uint32(tb.masked_crc32c(n)).tobytes() + payload + np.uint32(tb.masked_crc32c(payload)).tobytes())


def test_reader_parses_protobuf_written_examples_packed_and_unpacked(tmp_path):
    """Records serialised by protobuf itself (the bytes the reference's writer produces), plus a hand-made
    UNPACKED float / int64 encoding that proto2-era writers emit."""
    SequenceExample = _example_classes()
    rng = np.random.default_rng(2)
    x = rng.standard_normal((7, 6)).astype(np.float32)
    ex = SequenceExample()
    ex.context.feature['input_length'].int64_list.value.append(7)
    ex.context.feature['input_size'].int64_list.value.append(6)
    ex.context.feature['filename'].bytes_list.value.append(b's1')
    for row in x:
        ex.feature_lists.feature_list['inputs'].feature.add().float_list.value.extend(row.tolist())
    lab = SequenceExample()
    lab.context.feature['unit'].bytes_list.value.append(b'character')
    lab.context.feature['labels_length'].int64_list.value.append(3)
    lab.context.feature['filename'].bytes_list.value.append(b's1')
    for v in (7, 300, 2):
        lab.feature_lists.feature_list['labels'].feature.add().int64_list.value.append(v)

    def varint(v):
        out = b''
        while v >= 0x80:
            out += bytes([v & 0x7F | 0x80])
            v >>= 7
        return out + bytes([v])

    def ld(field, payload):
        return varint(field << 3 | 2) + varint(len(payload)) + payload

    def unpacked_floats(vals):  # Feature{float_list{value: wire type 5 each}}
        return ld(2, b''.join(b'\x0d' + np.float32(v).tobytes() for v in vals))

    y = rng.standard_normal((2, 6)).astype(np.float32)
    ctx = b''.join(ld(1, ld(1, k) + ld(2, v)) for k, v in (
        (b'input_length', ld(3, b'\x08\x02')),  # Int64List with an UNPACKED varint value
        (b'input_size', ld(3, ld(1, b'\x06'))),
        (b'filename', ld(1, ld(1, b's2')))))
    lists = ld(1, ld(1, b'inputs') + ld(2, b''.join(ld(1, unpacked_floats(r)) for r in y)))
    hand = ld(1, ctx) + ld(2, lists)
    p = str(tmp_path / 'a.tfrecord')
    open(p, 'wb').write(_frame(ex.SerializeToString()) + _frame(hand))
    open(str(tmp_path / 'l.tfrecord'), 'wb').write(_frame(lab.SerializeToString()))
    f = tfrecord.RecordFile(p, verify_data=True)
    assert (f.kind, f.n, f.feat, f.input_shape, f.has_aus) == (tfrecord.KIND_FEATURE, 2, 6, [6], False)
    assert f.lengths.tolist() == [7, 2] and f.filename(0) == b's1' and f.filename(1) == b's2'
    dst = np.full((2, 8, 6), np.nan, np.float32)
    lens = np.zeros(2, np.int32)
    f.fill_inputs([0, 1], 8, dst, lens, n_threads=2)
    assert lens.tolist() == [7, 2]
    assert np.array_equal(dst[0, :7], x) and np.all(dst[0, 7:] == 0)
    assert np.array_equal(dst[1, :2], y) and np.all(dst[1, 2:] == 0)
    f.fill_inputs([0], 8, dst[:1], lens[:1], reverse=True)
    assert np.array_equal(dst[0, :7], x[::-1]) and np.all(dst[0, 7:] == 0)
    lf = tfrecord.RecordFile(str(tmp_path / 'l.tfrecord'))
    assert lf.kind == tfrecord.KIND_LABELS and lf.unit == 'character'
    ids, ll = np.zeros((1, 5), np.int32), np.zeros(1, np.int32)
    lf.fill_labels([0], 5, 29, ids, ll)
    assert ids.tolist() == [[7, 300, 2, 29, 0]] and ll.tolist() == [4]  # EOS appended, io_utils.py:80-83


def test_corruption_is_detected(tmp_path):
    p = str(tmp_path / 'x.tfrecord')
    with tfrecord.RecordWriter(p) as w:
        w.write_feature('a', np.ones((3, 4), np.float32))
        w.write_feature('b', np.ones((2, 4), np.float32))
    raw = bytearray(open(p, 'rb').read())
    bad = bytearray(raw)
    bad[3] ^= 0x40  # length word
    open(p, 'wb').write(bad)
    with pytest.raises(tfrecord.AvsrIoError, match='length'):
        tfrecord.RecordFile(p)
    bad = bytearray(raw)
    bad[40] ^= 0x01  # payload
    open(p, 'wb').write(bad)
    with pytest.raises(tfrecord.AvsrIoError, match='data'):
        tfrecord.RecordFile(p, verify_data=True)
    open(p, 'wb').write(raw[:-7])
    with pytest.raises(tfrecord.AvsrIoError, match='truncated'):
        tfrecord.RecordFile(p)
    open(p, 'wb').write(b'')
    assert len(tfrecord.RecordFile(p)) == 0
    with pytest.raises(tfrecord.AvsrIoError):
        tfrecord.RecordFile(str(tmp_path / 'missing.tfrecord'))


def _write_set(tmp_path, n, seed=0, aus=False, Tv=(3, 20), ratio=4):
    rng = np.random.default_rng(seed)
    vp, ap, lp = (str(tmp_path / f'{k}.tfrecord') for k in ('video', 'audio', 'labels'))
    data = []
    with tfrecord.RecordWriter(vp) as wv, tfrecord.RecordWriter(ap) as wa, tfrecord.RecordWriter(lp) as wl:
        for i in range(n):
            tv = int(rng.integers(*Tv))
            v = rng.uniform(-1, 1, (tv, 4, 4, 3)).astype(np.float32)
            a = rng.standard_normal((tv * ratio, 5)).astype(np.float32)
            y = rng.integers(1, 29, int(rng.integers(2, 9)))
            au = rng.uniform(0, 4, (tv, 2)).astype(np.float32) if aus else None
            sid = 'utt%06d' % i
            wv.write_video(sid, v, au)
            wa.write_feature(sid, a)
            wl.write_labels(sid, y)
            data.append((v, a, y, au))
    return vp, ap, lp, data


def test_two_record_iterator_buckets_pads_and_keeps_streams_aligned(tmp_path):
    unit_dict = create_unit_dict(None)
    vp, ap, lp, data = _write_set(tmp_path, 57, aus=True)
    it = io_utils.make_iterator_from_two_records(vp, ap, lp, batch_size=8, unit_dict=unit_dict, shuffle=True,
                                                 bucket_width=5, seed=3, shuffle_buffer=16)
    seen = []
    for epoch in range(2):
        it.iterator_initializer()
        names = []
        while True:
            try:
                it.next()
            except io_utils.OutOfRangeError:
                break
            (v, a), (vl, al) = it.inputs, it.inputs_length
            B = v.shape[0]
            assert B <= 8 and v.shape[2:] == (4, 4, 3) and a.shape[2] == 5
            assert len({int(l) // 5 for l in vl}) == 1  # one bucket per batch (group_by_window key)
            assert v.shape[1] == int(vl.max()) and a.shape[1] == int(al.max())
            for b in range(B):
                i = int(it.labels_filenames[b][3:])
                vv, aa, yy, au = data[i]
                assert it.inputs_filenames[0][b] == it.inputs_filenames[1][b] == it.labels_filenames[b]
                assert int(vl[b]) == len(vv) and int(al[b]) == len(aa)
                assert np.array_equal(v[b, :len(vv)].numpy(), vv) and not v[b, len(vv):].any()
                assert np.array_equal(a[b, :len(aa)].numpy(), aa) and not a[b, len(aa):].any()
                assert np.array_equal(it.payload['aus'][b, :len(vv)].numpy(), au)
                n = int(it.labels_length[b])
                assert n == len(yy) + 1 and it.labels[b, :n].tolist() == list(yy) + [29]
                assert not it.labels[b, n:].any()
            vid, aud = it.data_sequences()
            assert vid.inputs is v and aud.inputs is a and 'aus' in vid.payload and aud.payload == {}
            names.extend(it.labels_filenames.tolist())
        assert sorted(names) == sorted(b'utt%06d' % i for i in range(57))  # every utterance exactly once
        seen.append(names)
    assert seen[0] != seen[1]  # reshuffle_each_iteration
    with pytest.raises(io_utils.OutOfRangeError):
        it.next()


def test_one_record_iterator_order_filter_and_reverse(tmp_path):
    unit_dict = create_unit_dict(None)
    vp, ap, lp, data = _write_set(tmp_path, 21)
    it = io_utils.make_iterator_from_one_record(ap, lp, unit_dict, batch_size=4, shuffle=False, prefetch=0)
    order = [n for b in it for n in b.labels_filenames.tolist()]
    assert order == [b'utt%06d' % i for i in range(21)]  # no shuffle: file order, last batch partial
    sizes = [len(b.labels_filenames) for b in it]
    assert sizes == [4, 4, 4, 4, 4, 1]
    it = io_utils.make_iterator_from_one_record(ap, lp, unit_dict, batch_size=4, max_sentence_length=6, prefetch=0)
    kept = [n for b in it for n in b.labels_filenames.tolist()]
    assert kept == [b'utt%06d' % i for i, d in enumerate(data) if len(d[2]) + 1 < 6]
    it = io_utils.make_iterator_from_one_record(ap, lp, unit_dict, batch_size=3, reverse_input=True, prefetch=0)
    b = next(iter(it))
    for k in range(3):
        a = data[k][1]
        assert np.array_equal(b.inputs[k, :len(a)].numpy(), a[::-1])
    vid, aud = b.data_sequences()
    assert vid is None and aud.inputs is b.inputs
    lit = io_utils.make_iterator_from_label_record(lp, batch_size=5, unit_dict=unit_dict, prefetch=0)
    b = next(iter(lit))
    assert b.inputs is None and b.labels.shape[0] == 5


def test_data_parallel_shards_walk_the_same_batches(tmp_path):
    """shard=(rank, world): same global batches on every rank (same seed), disjoint contiguous slices, the same number
    of steps everywhere (batches smaller than the world are dropped), union = the unsharded epoch."""
    unit_dict = create_unit_dict(None)
    vp, ap, lp, data = _write_set(tmp_path, 45)
    kw = dict(batch_size=8, unit_dict=unit_dict, shuffle=True, bucket_width=4, seed=11, shuffle_buffer=8, prefetch=0)
    full = io_utils.make_iterator_from_two_records(vp, ap, lp, **kw)
    world = 3
    ranks = [io_utils.make_iterator_from_two_records(vp, ap, lp, shard=(r, world), **kw) for r in range(world)]
    glob = full.batches_of_epoch()
    per_rank = [it.batches_of_epoch() for it in ranks]
    assert len({len(b) for b in per_rank}) == 1  # same number of steps on every rank
    kept = [g for g in glob if len(g) >= world]
    assert len(per_rank[0]) == len(kept)
    for step, g in enumerate(kept):
        parts = [per_rank[r][step] for r in range(world)]
        assert np.array_equal(np.concatenate(parts), g)
        sizes = [len(p) for p in parts]
        assert max(sizes) - min(sizes) <= 1 and min(sizes) >= 1
    # the assembled batch of a rank is its slice of the global one (a fresh epoch reshuffles: rewind the generator)
    ranks[1]._rng = np.random.default_rng(11)
    first = ranks[1].batches_of_epoch()[0]
    ranks[1]._rng = np.random.default_rng(11)
    b = next(iter(ranks[1]))
    assert b.labels_filenames.tolist() == [b'utt%06d' % i for i in first]


def test_golden_records_written_by_protobuf():
    """tests/golden/sequence_examples_*.tfrecord were serialised by google.protobuf and framed with tensorboard's crc
    (make_golden_records.py); the native reader must decode them to the JSON stored beside them."""
    import json
    g = os.path.join(ROOT, 'tests', 'golden')
    want = json.load(open(os.path.join(g, 'sequence_examples.json')))
    f = tfrecord.RecordFile(os.path.join(g, 'sequence_examples_feature.tfrecord'), verify_data=True)
    v = tfrecord.RecordFile(os.path.join(g, 'sequence_examples_video.tfrecord'), verify_data=True)
    l = tfrecord.RecordFile(os.path.join(g, 'sequence_examples_labels.tfrecord'), verify_data=True)
    assert (f.kind, f.input_shape, f.has_aus) == (tfrecord.KIND_FEATURE, [5], False)
    assert (v.kind, v.input_shape, v.has_aus) == (tfrecord.KIND_VIDEO, [2, 3, 3], True)  # [width, height, channels]
    assert (l.kind, l.unit) == (tfrecord.KIND_LABELS, 'character')
    n = len(want['labels'])
    assert len(f) == len(v) == len(l) == n
    idx = np.arange(n)
    tf_, tv = int(f.lengths.max()), int(v.lengths.max())
    xf, lf = np.zeros((n, tf_, 5), np.float32), np.zeros(n, np.int32)
    xv, av, lv = np.zeros((n, tv, 18), np.float32), np.zeros((n, tv, 2), np.float32), np.zeros(n, np.int32)
    f.fill_inputs(idx, tf_, xf, lf)
    v.fill_inputs(idx, tv, xv, lv, aus_dst=av)
    lp = int(l.lengths.max()) + 1
    y, ly = np.zeros((n, lp), np.int32), np.zeros(n, np.int32)
    l.fill_labels(idx, lp, 29, y, ly)
    for i in range(n):
        wf, wv, wl = want['feature'][i], want['video'][i], want['labels'][i]
        assert f.filename(i).decode() == wf['filename'] == v.filename(i).decode() == l.filename(i).decode()
        assert np.array_equal(xf[i, :lf[i]], np.asarray(wf['inputs'], np.float32)) and not xf[i, lf[i]:].any()
        assert np.array_equal(xv[i, :lv[i]], np.asarray(wv['inputs'], np.float32))
        assert np.array_equal(av[i, :lv[i]], np.asarray(wv['aus'], np.float32))
        assert y[i, :ly[i]].tolist() == wl['labels'] + [29]


def test_record_inspection_helpers_keep_the_reference_names():
    g = os.path.join(ROOT, 'tests', 'golden')
    assert io_utils._get_input_shape_from_record(os.path.join(g, 'sequence_examples_feature.tfrecord')) == \
        ([5], {'stream': 'feature'})
    assert io_utils._get_input_shape_from_record(os.path.join(g, 'sequence_examples_video.tfrecord')) == \
        ([2, 3, 3], {'stream': 'video', 'aus': True})
    assert io_utils._get_unit_from_record(os.path.join(g, 'sequence_examples_labels.tfrecord')) == 'character'
