"""Host-side logic that needs no GPU: the variable table (names / shapes / counts of SURVEY.md 8e),
hparams defaults, the reference's configuration errors, checkpoint round trip."""
import numpy as np
import pytest

from avsr_tf1_b200 import make_hparams
from avsr_tf1_b200.seq2seq import Seq2SeqModel
from tests.helpers import config_hparams, synthetic_batch, to_data_sequences


def build(cfg, **over):
    hp = config_hparams(cfg, **over)
    batch = synthetic_batch(hp, B=2, Ta=6, Tv=4, L=3)
    return hp, Seq2SeqModel(to_data_sequences(batch), 'train', hp, device='cpu')


@pytest.mark.parametrize('cfg,over,count', [
    (1, {}, 361408),
    (2, {}, 4639807),
    (3, dict(attention_type=(('bahdanau',), ('bahdanau',))), 2375839),
    (3, {}, 2310048),                                                             # scaled-Luong default
    (4, dict(attention_type=(('bahdanau',), ('bahdanau',))), 4427327),
    (5, dict(attention_type=(('bahdanau',), ('bahdanau',))), 4296255),
    (5, {}, 4164673),
])
def test_trainable_parameter_counts_match_survey(cfg, over, count):
    _, m = build(cfg, **over)
    assert m.n_params == count


def test_variable_names_follow_tf_scopes():
    _, m = build(5, attention_type=(('bahdanau',), ('scaled_luong',)))
    names = set(m.store.names(trainable_only=False))
    for n in ['video/batch_normalization/gamma', 'video/batch_normalization/moving_variance',
              'video/Encoder/multi_rnn_cell/cell_0/lstm_cell/kernel',
              'audio/Encoder/multi_rnn_cell/cell_2/attention_wrapper/lstm_cell/kernel',
              'audio/Encoder/multi_rnn_cell/cell_2/attention_wrapper/bahdanau_attention/query_layer/kernel',
              'audio/Encoder/multi_rnn_cell/cell_2/attention_wrapper/bahdanau_attention/attention_v',
              'audio/Encoder/multi_rnn_cell/cell_2/attention_wrapper/attention_layer/kernel',
              'audio/Encoder/memory_layer/kernel', 'embeddings/embedding_matrix', 'Decoder/memory_layer/kernel',
              'Decoder/decoder/attention_wrapper/lstm_cell/kernel',
              'Decoder/decoder/attention_wrapper/luong_attention/attention_g',
              'Decoder/decoder/attention_wrapper/attention_layer/kernel', 'Decoder/decoder/my_dense/kernel',
              'Decoder/decoder/my_dense/bias']:
        assert n in names, n
    assert m.store.p('audio/Encoder/multi_rnn_cell/cell_2/attention_wrapper/lstm_cell/kernel').shape == (768, 1024)
    assert m.store.p('Decoder/decoder/attention_wrapper/lstm_cell/kernel').shape == (128 + 256 + 256, 1024)
    _, b = build(4)
    assert b.store.p('Decoder/state_projection/kernel').shape == (512, 256)
    assert b.store.p('Decoder/decoder/attention_wrapper/lstm_cell/kernel').shape == (128 + 512 + 256, 1024)
    _, bi = build(2)
    assert bi.store.p('audio/Encoder/dense_5/kernel').shape == (512, 256)
    assert 'audio/Encoder/fw/multi_rnn_cell/cell_1/lstm_cell/kernel' in set(bi.store.names())
    # L2 filter of seq2seq.py:283-290 selects exactly the LSTM kernels
    assert sorted(bi._l2_names) == sorted(n for n in bi.store.names() if n.endswith('lstm_cell/kernel'))


def test_initialisers():
    _, m = build(1)
    P = m.store.to_numpy('p')
    k = P['audio/Encoder/lstm_cell/kernel']
    assert abs(k.std() - np.sqrt(1.0 / k.shape[0])) < 0.1 * np.sqrt(1.0 / k.shape[0])
    assert np.abs(k).max() <= 2.0 * np.sqrt(1.0 / k.shape[0]) / .87962566103423978 + 1e-6  # truncated normal
    assert np.all(P['audio/Encoder/lstm_cell/bias'] == 0)
    e = P['embeddings/embedding_matrix']
    assert e.shape == (31, 128) and np.abs(e).max() <= 1.732 / 31
    assert float(P['Decoder/decoder/attention_wrapper/luong_attention/attention_g'].reshape(-1)[0]) == 1.0
    assert np.all(P['audio/batch_normalization/moving_variance'] == 1)


def test_hparams_defaults_follow_reference():
    hp = make_hparams()
    assert hp.attention_type == (('scaled_luong',), ('scaled_luong',))       # avsr.py:50
    assert hp.encoder_units_per_layer == ((256,), (256, 256, 256))           # avsr.py:46
    assert hp.max_label_length == 150 and hp.beam_width == 10 and hp.embedding_size == 128
    assert hp.recurrent_l2_regularisation == 0.0001 and hp.max_gradient_norm == 1.0
    assert make_hparams(optimiser='AdamW').recurrent_l2_regularisation is None  # avsr.py:172
    assert make_hparams(warmup_steps=10).kwargs == {'warmup_steps': 10}


@pytest.mark.parametrize('over,exc', [
    (dict(architecture='nonsense'), Exception),
    (dict(encoder_type='sideways'), Exception),
    (dict(cell_type='gru'), Exception),
    (dict(attention_type=(('cosine',), ('cosine',))), Exception),
    (dict(decoding_algorithm='viterbi'), Exception),
    (dict(optimiser='SGD'), Exception),
    (dict(loss_fun='hinge'), ValueError),
])
def test_configuration_errors(over, exc):
    with pytest.raises(exc):
        build(1, **over)


def test_bidirectional_single_layer_is_rejected_like_the_reference():
    with pytest.raises(ValueError):  # encoder.py:125-133 cannot index a 1-layer state
        build(1, encoder_type='bidirectional')


def test_attentive_encoder_is_unidirectional_only():
    with pytest.raises(Exception):
        build(5, encoder_type='bidirectional')


def test_checkpoint_round_trip(tmp_path):
    _, m = build(1)
    m._global_step = 17
    m.store.m.normal_()
    path = m.saver.save(None, str(tmp_path / 'checkpoint.ckp'), global_step=30)
    assert path.endswith('checkpoint.ckp-30')
    _, m2 = build(1)
    m2.store.flat.zero_()
    m2.saver.restore(None, path)
    assert m2.global_step == 17
    a, b = m.store.to_numpy('p'), m2.store.to_numpy('p')
    assert all(np.array_equal(a[k], b[k]) for k in a)
    assert all(np.array_equal(v, m2.store.to_numpy('m')[k]) for k, v in m.store.to_numpy('m').items())


def test_reference_default_randomness_builds_and_gets_its_own_streams():
    """use_dropout=True / sampling_probability_outputs=0.1 are the reference defaults (avsr.py:49-56): every
    DropoutWrapper-ed cell and the sampling helper get disjoint generator streams; evaluate mode gets none."""
    hp, m = build(5, use_dropout=True, sampling_probability_outputs=0.1)
    streams = m.random_streams
    assert 'Decoder/sampling' in streams
    cells = [n for n in streams if n != 'Decoder/sampling']
    assert len(cells) == 3 + 3 + 1  # video 3, audio 2 + cross-modal wrapper, decoder wrapper
    spans = sorted((s, s + (2 if n == 'Decoder/sampling' else 4)) for n, s in streams.items())
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))
    from avsr_tf1_b200.seq2seq import Seq2SeqModel
    from tests.helpers import synthetic_batch, to_data_sequences
    ds = to_data_sequences(synthetic_batch(hp, B=2, Ta=8, Tv=4, L=3))
    ev = Seq2SeqModel(ds, 'evaluate', hp, device='cpu')
    assert ev.random_streams == {}


def test_cosine_decay_restarts_schedule():
    """tf.train.cosine_decay_restarts(lr, step, first_decay_steps) with t_mul = 2, m_mul = 1, alpha = 0: hand values of
    the closed form, then the model's lr with the 750-step warm-up on top (seq2seq.py:263-280)."""
    from avsr_tf1_b200.seq2seq import cosine_decay_restarts as cdr
    lr, first = 1e-3, 100
    assert cdr(lr, 0, first) == pytest.approx(lr)
    assert cdr(lr, 50, first) == pytest.approx(0.5 * lr)
    assert cdr(lr, 99, first) == pytest.approx(0.5 * lr * (1 + np.cos(np.pi * 0.99)))
    assert cdr(lr, 100, first) == pytest.approx(lr)            # first restart, period 200
    assert cdr(lr, 200, first) == pytest.approx(0.5 * lr)      # halfway through it
    assert cdr(lr, 300, first) == pytest.approx(lr)            # second restart, period 400
    assert cdr(lr, 500, first) == pytest.approx(0.5 * lr)
    assert cdr(lr, 25, first, t_mul=1.0) == pytest.approx(cdr(lr, 125, first, t_mul=1.0))
    hp, m = build(1, lr_decay=('cosine_restarts', 100))
    m._global_step = 50
    assert m._lr_now() == pytest.approx(0.5 * hp.learning_rate * 51 / 750.0)
    hp, m = build(1, lr_decay=('exponential', 10))  # unknown policy: constant lr, like the reference
    assert m._lr_now() == pytest.approx(hp.learning_rate / 750.0)


def test_png_writer_round_trips(tmp_path):
    cv2 = pytest.importorskip('cv2')
    from avsr_tf1_b200.utils import write_png_gray
    img = np.random.default_rng(0).uniform(0, 1, (37, 53))
    f = str(tmp_path / 'a.png')
    write_png_gray(f, img)
    back = cv2.imread(f, cv2.IMREAD_GRAYSCALE)
    assert back.shape == (37, 53)
    assert np.array_equal(back, (img * 255.0 + 0.5).astype(np.uint8))


def test_resnet_cnn_variable_table():
    """video_processing='resnet_cnn' on 36x36x3 crops adds the 282 288 trainable CNN parameters of SURVEY.md 8e under
    the TF names of video.py (conv2d / batch_normalization `name=` arguments inside variable_scope('CNN'))."""
    from tests.helpers import to_image_sequences
    hp = config_hparams(3, video_processing='resnet_cnn', attention_type=(('bahdanau',), ('bahdanau',)))
    batch = to_image_sequences(synthetic_batch(hp, B=2, Ta=6, Tv=4, L=3), hw=36)
    m = Seq2SeqModel(to_data_sequences(batch), 'train', hp, device='cpu')
    assert m.n_params == 2375839 + 282288
    names = set(m.store.names(trainable_only=False))
    for n in ('CNN/layer0/kernel', 'CNN/layer0_bn/moving_variance', 'CNN/res_block_0_conv1/bias',
              'CNN/res_block_0_second_bn/gamma', 'CNN/res_block_1_first_bn/beta', 'CNN/res_block_3_shortcut/kernel',
              'CNN/res_block_3_conv2/kernel', 'CNN/flatten/kernel'):
        assert n in names, n
    assert 'CNN/res_block_0_first_bn/gamma' not in names and 'CNN/res_block_0_shortcut/kernel' not in names
    assert m.store.table['CNN/flatten/kernel'][1] == (5, 5, 64, 128)  # 36 -> 18 -> 9 -> 5, VALID over the rest
    assert m.store.table['CNN/res_block_2_shortcut/kernel'][1] == (1, 1, 16, 32)


def test_host_ring_is_bounded_and_reuses_in_order():
    """the staging ring keyed buffers by exact shape and never evicted (one set per padded length of a
    bucketed epoch).  Now: `depth` flat buffers per role, grown geometrically, handed out in slot order 0, 1, ..."""
    import torch
    from avsr_tf1_b200.io_utils import _HostRing
    ring = _HostRing(pin=False, depth=3)
    first = [ring.get(('x', 0), (4, 10, 8), torch.float32) for _ in range(3)]
    assert len({t.data_ptr() for t in first}) == 3
    again = ring.get(('x', 0), (4, 10, 8), torch.float32)
    assert again.data_ptr() == first[0].data_ptr()  # reuse distance = depth: the OLDEST slot comes back first
    for t_pad in range(11, 200):                    # hundreds of distinct padded lengths ...
        t = ring.get(('x', 0), (4, t_pad, 8), torch.float32)
        assert t.shape == (4, t_pad, 8) and t.is_contiguous()
    biggest = 4 * 199 * 8 * 4
    assert ring.pinned_bytes() <= 3 * int(1.5 * biggest)  # ... still three buffers, each at most 1.5x the largest batch
    a = ring.get(('aus', 0), (4, 7, 2), torch.float32)    # another role has its own ring
    a.fill_(1.0)
    assert float(a.sum()) == 4 * 7 * 2


def test_record_iterator_stays_failed_after_a_producer_error(tmp_path):
    """after the prefetch thread raised once, the next next() blocked forever on the dead thread's queue."""
    from avsr_tf1_b200 import io_utils
    from avsr_tf1_b200.synthetic import write_synthetic_records
    paths = write_synthetic_records(str(tmp_path), n=6, Ta=8, Tv=4, Fa=3, hw=2, L=3)
    from avsr_tf1_b200.hparams import create_unit_dict
    it = io_utils.make_iterator_from_one_record(paths["audio"], paths["labels"], create_unit_dict(None), batch_size=2,
                                                prefetch=1, pin_memory=False)

    def boom(idx):
        raise RuntimeError('decode failed')
    it._assemble = boom
    it.iterator_initializer()
    with pytest.raises(RuntimeError):
        it.next()
    with pytest.raises(io_utils.OutOfRangeError):  # does not hang: the iterator is at its end until re-initialised
        it.next()


def test_checkpoint_is_written_atomically(tmp_path):
    from avsr_tf1_b200.avsr import latest_checkpoint
    from avsr_tf1_b200.seq2seq import Seq2SeqModel
    hp = config_hparams(1)
    batch = synthetic_batch(hp, B=2, Ta=6, L=3)
    model = Seq2SeqModel(to_data_sequences(batch), 'train', hp, seed=1, device='cpu')
    ckp = str(tmp_path / 'checkpoint.ckp')
    model.saver.save(None, ckp, global_step=10)
    # a crash during the NEXT save leaves only a temporary file behind: it is never taken for a checkpoint
    open(ckp + '-20.tmp.npz', 'wb').write(b'truncated')
    assert latest_checkpoint(str(tmp_path)) == ckp + '-10'
    model.saver.save(None, ckp, global_step=20)
    assert latest_checkpoint(str(tmp_path)) == ckp + '-20'
    assert not (tmp_path / 'checkpoint.ckp-10.npz').exists()  # max_to_keep = 1, removed only after the rename
    assert not (tmp_path / 'checkpoint.ckp-20.tmp.npz').exists()


def test_variable_tables_of_the_round_2_options():
    """Names / shapes the non-default options add or remove (TF scopes of the reference graph)."""
    def table(cfg, **over):
        _, m = build(cfg, **over)
        return {n: tuple(m.store.p(n).shape) for n in m.store.names()}
    # HighwayWrapper: carry variables in the position's own scope, per position even with the shared cell (cells.py:77-92)
    t = table(3, highway_encoder=True, encoder_weight_sharing=True)
    for k in (1, 2):
        assert t[f'video/Encoder/multi_rnn_cell/cell_{k}/carry_w'] == (256, 256)
        assert t[f'video/Encoder/multi_rnn_cell/cell_{k}/carry_b'] == (256,)
    assert 'video/Encoder/multi_rnn_cell/cell_0/carry_w' not in t
    assert 'video/Encoder/multi_rnn_cell/cell_2/lstm_cell/kernel' not in t  # layers 2.. run with the cell of layer 1
    # enable_attention=False: the bare cell lives in the decoder scope; no mechanism, no memory layer
    t = table(1, enable_attention=False)
    assert t['Decoder/decoder/lstm_cell/kernel'] == (128 + 128, 4 * 128)
    assert not any('attention' in n or 'memory_layer' in n for n in t)
    # one-hot decoder inputs: tf.eye is a constant, the cell input is the alphabet
    t = table(1, embedding_size=0)
    assert 'embeddings/embedding_matrix' not in t
    assert t['Decoder/decoder/attention_wrapper/lstm_cell/kernel'] == (31 + 128 + 128, 4 * 128)
    # instance_norm after the batch norm
    t = table(5, instance_normalisation=True)
    assert t['audio/InstanceNorm/gamma'] == (80,) and t['video/InstanceNorm/beta'] == (128,)
    assert t['audio/batch_normalization/gamma'] == (80,)
    # bimodal decoder with the video stream missing: one mechanism, zero state of the FIRST layer's width in the projection
    t = table(4, video_processing=None)
    assert t['Decoder/state_projection/kernel'] == (256 + 256, 256)
    assert 'Decoder/memory_layer_1/kernel' not in t and t['Decoder/memory_layer/kernel'] == (256, 256)
    assert t['Decoder/decoder/attention_wrapper/lstm_cell/kernel'] == (128 + 256 + 256, 4 * 256)
    assert not any(n.startswith('video/') for n in t)
