"""The runtime shell end to end on the GPU: synthetic TFRecords -> AVSR.train (epoch loop, logfile,
checkpoint + evaluation at epoch 10, .mlf dump) -> resume from the checkpoint -> evaluate, in the reference's call
pattern (avsr.py:227-512, experiment.py)."""
import os
import re

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def records(tmp_path_factory):
    from avsr_tf1_b200.synthetic import write_synthetic_records
    d = tmp_path_factory.mktemp('records')
    train = write_synthetic_records(str(d), n=24, Ta=32, Tv=8, Fa=20, hw=6, L=6, ragged=True, with_aus=True,
                                    prefix='train')
    test = write_synthetic_records(str(d), n=10, Ta=32, Tv=8, Fa=20, hw=6, L=6, ragged=True, with_aus=True,
                                   prefix='test', seed=50)
    return train, test


def make(records, workdir, **over):
    from avsr_tf1_b200.avsr import AVSR
    train, test = records
    kw = dict(unit='character', video_processing='features', audio_processing='features',
              video_train_record=train['video'], video_test_record=test['video'],
              audio_train_record=train['audio'], audio_test_record=test['audio'],
              labels_train_record=train['labels'], labels_test_record=test['labels'],
              batch_size=(8, 4), architecture='av_align', encoder_units_per_layer=((256,), (256, 256)),
              decoder_units_per_layer=(256,), decoding_algorithm='greedy', regress_aus=True, workdir=str(workdir),
              verbose=False)
    kw.update(over)
    return AVSR(**kw)


def test_train_checkpoint_evaluate_resume(records, tmp_path):
    logfile = str(tmp_path / 'logs' / 'run1')
    exp = make(records, tmp_path)  # reference defaults: dropout 0.9 and scheduled sampling 0.1 are ON
    exp.train(logfile=logfile, num_epochs=11)
    lines = open(logfile).read().splitlines()
    losses = [float(m.group(2)) for m in (re.match(r'Average batch_loss as epoch (\d+) is (\S+)', l) for l in lines) if m]
    assert len(losses) == 10 and all(np.isfinite(losses))
    assert losses[-1] < losses[0]  # it learns something in 10 epochs
    assert re.match(r'character: \d+\.\d{4}% word: \d+\.\d{4}% ', lines[-1])
    ckpt = os.path.join(str(tmp_path), 'checkpoints', 'run1', 'checkpoint.ckp-10')
    assert os.path.exists(ckpt + '.npz')
    mlf = os.path.join(str(tmp_path), 'predictions', 'run1', 'predicted_epoch_10.mlf')
    rows = open(mlf).read().splitlines()
    assert len(rows) == 10 and all(re.match(r'utt\d{6} .*\[.*\] \[\d+\.\d{3}\]$', r) for r in rows)
    steps = exp._train_model.model.global_step
    assert steps == 10 * 3  # 24 utterances / batch 8, bucketed (one bucket at these lengths)

    # a fresh object resumes from the latest checkpoint: epochs continue at 11, global_step survives (warm-up once)
    exp2 = make(records, tmp_path, learning_rate=0.0001)
    exp2.train(logfile=logfile, num_epochs=2, try_restore_latest_checkpoint=True)
    lines = open(logfile).read().splitlines()
    assert lines[-1].startswith('Average batch_loss as epoch 11 is ')
    assert exp2._train_model.model.global_step == steps + 3
    # evaluate() alone, from the checkpoint, reproduces the logged error rate (greedy decoding is deterministic)
    ev = make(records, tmp_path, required_grahps=('eval',))
    rate = ev.evaluate(ckpt, epoch=10)
    assert set(rate) == {'character', 'word'}
    assert abs(rate['character'] - exp.last_error_rate['character']) < 1e-9


def test_run_experiment_curriculum(records, tmp_path):
    from avsr_tf1_b200.experiment import run_experiment
    train, test = records
    run_experiment(labels_train_record=train['labels'], labels_test_record=test['labels'],
                   audio_train_records=(train['audio'],), audio_test_records=(test['audio'],),
                   iterations=((2, 1),), learning_rates=((0.001, 0.0001),), architecture='unimodal',
                   logfile='exp', logdir=str(tmp_path / 'logs'), audio_processing='features',
                   encoder_units_per_layer=((128,), (128,)), decoder_units_per_layer=(128,), batch_size=(8, 4),
                   warmup_epochs=2, warmup_max_len=6, workdir=str(tmp_path), verbose=False,
                   decoding_algorithm='greedy')
    lines = open(str(tmp_path / 'logs' / 'exp')).read().splitlines()
    assert lines[0].startswith('Warm up on short sentences up to 6 tokens for 2 epochs')
    assert lines.count('=====') == 2 and lines.count('=' * 20) == 1
    assert sum(l.startswith('Average batch_loss') for l in lines) == 1 + 2 + 1


def test_alignment_images(records, tmp_path):
    """write_attention_alignment=True (avsr.py:353-436): alignment_history of the decoder and of the cross-modal encoder
    in the reference's [B, T_memory, T_query, 1] layout, dumped as <file>.png / <file>_av.png (pixel = 1 - weight)."""
    cv2 = pytest.importorskip('cv2')
    exp = make(records, tmp_path, write_attention_alignment=True, use_dropout=False, sampling_probability_outputs=0.0)
    exp.train(logfile=str(tmp_path / 'logs' / 'al'), num_epochs=2)
    ckpt = exp._train_model.model.saver.save(None, str(tmp_path / 'checkpoints' / 'al' / 'checkpoint.ckp'), global_step=1)
    out = str(tmp_path / 'alignments')
    exp.evaluate(ckpt, epoch=1, alignments_outdir=out)
    model = exp._evaluate_model.model
    dec, enc = model._decoder.attention_alignment, model._audio_encoder.attention_alignment
    B = dec.shape[0]
    assert dec.ndim == 4 and dec.shape[3] == 1 and enc.shape[0] == B and enc.shape[3] == 1
    assert dec.shape[1] == enc.shape[2]  # decoder memory = audio steps; encoder memory = video steps
    ids = model._decoder.inference_predicted_ids
    for b in range(B):
        steps = int((ids[b] != 0).sum())  # steps before the row finished carry a distribution over the memory
        col = dec[b, :, :steps, 0].sum(axis=0)
        assert np.allclose(col, 1.0, atol=1e-4), col
    pngs = sorted(os.listdir(out))
    assert len(pngs) == 2 * 10 and pngs[0] == 'utt000000.png' and pngs[1] == 'utt000000_av.png'
    img = cv2.imread(os.path.join(out, pngs[-2]), cv2.IMREAD_GRAYSCALE)
    assert img.shape == dec.shape[1:3]  # last batch: same padded sizes as the arrays still held by the model


def test_runtime_shell_with_cnn_front_end(tmp_path):
    """run_video.py's configuration in miniature: AVSR(video_processing='resnet_cnn') on a video record of lip crops."""
    from avsr_tf1_b200.avsr import AVSR
    from avsr_tf1_b200.synthetic import write_synthetic_records
    train = write_synthetic_records(str(tmp_path), n=12, Tv=6, hw=12, L=4, ragged=True, audio=False, prefix='tr')
    test = write_synthetic_records(str(tmp_path), n=5, Tv=6, hw=12, L=4, ragged=True, audio=False, prefix='te', seed=9)
    exp = AVSR(unit='character', video_processing='resnet_cnn', video_train_record=train['video'],
               video_test_record=test['video'], labels_train_record=train['labels'],
               labels_test_record=test['labels'], batch_size=(6, 5), architecture='unimodal',
               encoder_units_per_layer=((128,), (128,)), decoder_units_per_layer=(128,), cnn_filters=(4, 8, 8, 16),
               cnn_dense_units=32, decoding_algorithm='greedy', workdir=str(tmp_path), verbose=False)
    exp.train(logfile=str(tmp_path / 'logs' / 'cnn'), num_epochs=11)
    lines = open(str(tmp_path / 'logs' / 'cnn')).read().splitlines()
    assert sum(l.startswith('Average batch_loss') for l in lines) == 10 and lines[-1].startswith('character: ')
    assert any(n.startswith('CNN/') for n in exp._train_model.model.store.names())
