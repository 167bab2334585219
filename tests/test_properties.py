"""Property tests (hypothesis) of the integer / host-side pieces: the edit distance behind CER / WER (bit-exact work:
product and oracle must agree on everything, and the metric axioms hold), the beam back-trace, the TFRecord round trip
of arbitrary examples, the batching iterator's coverage."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from avsr_tf1_b200 import tfrecord, utils
from oracle import avsr_oracle as O

tokens = st.lists(st.sampled_from(list('abc d') + ["'"]), max_size=12)


@settings(max_examples=200, deadline=None)
@given(tokens, tokens, tokens)
def test_levenshtein_is_a_metric_and_matches_the_oracle(a, b, c):
    d = utils.levenshtein
    assert d(a, b) == O.levenshtein(a, b) == d(b, a)
    assert d(a, a) == 0 and (d(a, b) == 0) == (a == b)
    assert d(a, c) <= d(a, b) + d(b, c)
    assert abs(len(a) - len(b)) <= d(a, b) <= max(len(a), len(b))


@settings(max_examples=100, deadline=None)
@given(st.integers(1, 6), st.integers(1, 3), st.integers(1, 4), st.data())
def test_gather_tree_follows_parents_and_pads_after_eos(T, B, W, data):
    from avsr_tf1_b200.decoder_unimodal import gather_tree
    eos = 9
    step = np.array(data.draw(st.lists(st.integers(0, 9), min_size=T * B * W, max_size=T * B * W))).reshape(T, B, W)
    parent = np.array(data.draw(st.lists(st.integers(0, W - 1), min_size=T * B * W, max_size=T * B * W))).reshape(T, B, W)
    max_len = np.array(data.draw(st.lists(st.integers(0, T), min_size=B, max_size=B)))
    out = gather_tree(step, parent, max_len, eos)
    for b in range(B):
        ml = int(max_len[b])
        assert (out[ml:, b] == eos).all()
        for w in range(W):
            # independent walk from the last valid step back to the first
            seq, p = [], w
            for t in range(ml - 1, -1, -1):
                seq.append(step[t, b, p])
                p = parent[t, b, p]
            seq = seq[::-1]
            if eos in seq:
                k = seq.index(eos)
                seq = seq[:k] + [eos] * (ml - k)
            assert out[:ml, b, w].tolist() == seq


@settings(max_examples=40, deadline=None)
@given(st.integers(0, 7), st.integers(1, 6), st.data())
def test_tfrecord_round_trip_of_arbitrary_examples(steps, size, data):
    import os
    import tempfile
    vals = data.draw(st.lists(st.floats(-1e6, 1e6, width=32), min_size=steps * size, max_size=steps * size))
    labels = data.draw(st.lists(st.integers(0, 2 ** 31 - 1), max_size=9))  # stored as int64 varints, read back as int32
    name = data.draw(st.text(alphabet='abcXYZ019_/-.', max_size=20))
    x = np.array(vals, np.float32).reshape(steps, size)
    with tempfile.TemporaryDirectory() as d:
        pf, pl = os.path.join(d, 'f.tfrecord'), os.path.join(d, 'l.tfrecord')
        with tfrecord.RecordWriter(pf) as w:
            w.write_feature(name, x)
        with tfrecord.RecordWriter(pl) as w:
            w.write_labels(name, labels)
        f, l = tfrecord.RecordFile(pf, verify_data=True), tfrecord.RecordFile(pl, verify_data=True)
        assert f.lengths.tolist() == [steps] and f.filename(0) == name.encode() == l.filename(0)
        dst, lens = np.full((1, steps + 2, size), 7, np.float32), np.zeros(1, np.int32)
        f.fill_inputs([0], steps + 2, dst, lens)
        assert lens[0] == steps and np.array_equal(dst[0, :steps], x) and not dst[0, steps:].any()
        ids, ll = np.zeros((1, len(labels) + 1), np.int32), np.zeros(1, np.int32)
        l.fill_labels([0], len(labels) + 1, 29, ids, ll)
        assert ll[0] == len(labels) + 1 and ids[0, -1] == 29
        assert ids[0, :-1].tolist() == labels
        f.close()
        l.close()
