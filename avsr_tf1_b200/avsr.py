"""AVSR - the runtime shell around Seq2SeqModel, drop-in for reference avsr/avsr.py (class AVSR :19-760;
SURVEY.md section 8, row f-1): epoch loop, logfile lines, checkpoint every 10 epochs followed by an evaluation,
predictions dumped as .mlf, error rates via utils.compute_wer.

What differs by construction: there is no TF graph / session.  The train and evaluate "graphs" are two
Seq2SeqModel objects fed by eager record iterators (io_utils.RecordBatcher); a checkpoint is the Saver's .npz.
Video enters as the features a record holds (`features`, also raw lip crops as flat vectors) or as lip crops through
the `resnet_cnn` front-end (video.py); the other CNNs and `wav` audio raise."""
from __future__ import annotations

import collections
import glob
import re
import time
from os import devnull, makedirs, path

import numpy as np

from . import parallel
from .hparams import create_unit_dict, make_hparams
from .io_utils import (BatchedData, OutOfRangeError, make_iterator_from_one_record, make_iterator_from_two_records)
from .utils import compute_wer, write_sequences_to_labelfile


class Model(collections.namedtuple("Model", ("data", "model", "initializer", "batch_size"))):
    pass


def latest_checkpoint(checkpoint_dir):
    """tf.train.latest_checkpoint: the save path (without extension) with the highest epoch suffix, or None."""
    best, best_epoch = None, -1
    for f in glob.glob(path.join(checkpoint_dir, '*.npz')):
        m = re.search(r'-(\d+)\.npz$', f)
        if m and int(m.group(1)) > best_epoch:
            best, best_epoch = f[:-4], int(m.group(1))
    return best


class AVSR(object):
    def __init__(self, unit, unit_file=None, video_processing=None, video_train_record=None, video_test_record=None,
                 audio_processing=None, audio_train_record=None, audio_test_record=None, labels_train_record=None,
                 labels_test_record=None, batch_size=(64, 64), write_attention_alignment=False,
                 write_beam_search_graphs=False, write_estimated_modality_lags=False,
                 required_grahps=('train', 'eval'), workdir='.', seed=2001, verbose=True, **kwargs):
        """Keyword surface of avsr.py:21-75 (`required_grahps` is the reference's spelling); everything that is a
        hyper-parameter goes to make_hparams.  `workdir` roots the reference's relative output directories
        (checkpoints/, predictions/)."""
        if video_processing not in (None, 'features', 'resnet_cnn'):
            raise NotImplementedError('of the CNN front-ends (avsr/video.py) only `resnet_cnn` is built')
        if audio_processing is not None and audio_processing != 'features':
            raise NotImplementedError('`wav` audio processing (avsr/audio.py) is out of scope: use `features`')
        if write_beam_search_graphs or write_estimated_modality_lags:
            raise NotImplementedError('beam-search html graphs and modality-lag artefacts (avsr/visualise) are out '
                                      'of scope')
        self._write_attention_alignment = write_attention_alignment
        self._unit = unit
        self._unit_dict = create_unit_dict(unit_file=unit_file)
        self._video_processing, self._audio_processing = video_processing, audio_processing
        self._video_train_record, self._video_test_record = video_train_record, video_test_record
        self._audio_train_record, self._audio_test_record = audio_train_record, audio_test_record
        self._labels_train_record, self._labels_test_record = labels_train_record, labels_test_record
        self._required_graphs = required_grahps
        self._workdir, self._seed, self._verbose = workdir, seed, verbose
        self._hparams = make_hparams(unit=unit, unit_file=unit_file, video_processing=video_processing,
                                     audio_processing=audio_processing, batch_size=batch_size,
                                     unit_dict=self._unit_dict,
                                     write_attention_alignment=write_attention_alignment, **kwargs)
        self._train_model = self._evaluate_model = None
        self._create_models()

    # ---- construction (avsr.py:514-572) --------------------------------------------------------------------
    def _create_models(self):
        if 'train' in self._required_graphs:
            self._train_model = self._make_model('train', self._hparams.batch_size[0])
        if 'eval' in self._required_graphs:
            self._evaluate_model = self._make_model('evaluate', self._hparams.batch_size[1])

    def _fetch_data(self, mode, batch_size):
        """avsr.py:628-679: one iterator over both streams when both are on (bucket width 45 on the video length),
        else one iterator per stream."""
        train = mode == 'train'
        pick = (lambda a, b: a if train else b)
        labels = pick(self._labels_train_record, self._labels_test_record)
        video = pick(self._video_train_record, self._video_test_record)
        audio = pick(self._audio_train_record, self._audio_test_record)
        common = dict(batch_size=batch_size, unit_dict=self._hparams.unit_dict, shuffle=train, reverse_input=False,
                      bucket_width=45, seed=self._seed)
        if train:  # one process per GPU: batch_size is the GLOBAL batch, every rank trains on its slice of each batch
            common['shard'] = (parallel.rank(), parallel.world_size())
        if self._video_processing is not None and self._audio_processing is not None:
            return make_iterator_from_two_records(video_record=video, audio_record=audio, label_record=labels, **common)
        if self._video_processing is not None:
            return make_iterator_from_one_record(data_record=video, label_record=labels, **common)
        if self._audio_processing is not None:
            return make_iterator_from_one_record(data_record=audio, label_record=labels,
                                                 max_sentence_length=self._hparams.max_sentence_length, **common)
        raise ValueError('At least one of A/V streams must be enabled')

    def _make_model(self, mode, batch_size):
        from .seq2seq import Seq2SeqModel
        iterator = self._fetch_data(mode, batch_size)
        # the model only needs the feature sizes of each stream at construction: a one-step placeholder batch
        spec = []
        for f in iterator._inputs:
            cnn = self._video_processing is not None and 'cnn' in self._video_processing and len(f.input_shape) == 3
            # (the record stores [width, height, channels]; frames are laid out rows first: avsr/io_utils.py:318-332)
            x = np.zeros((1, 1) + ((f.input_shape[1], f.input_shape[0], f.input_shape[2]) if cnn else (f.feat,)),
                         np.float32)
            spec.append(BatchedData(iterator_initializer=iterator.iterator_initializer, inputs=x,
                                    inputs_length=np.ones(1, np.int32), inputs_filenames=None,
                                    labels=np.zeros((1, 1), np.int32), labels_length=np.ones(1, np.int32),
                                    labels_filenames=None, payload={}))
        if len(spec) == 2:
            data = (spec[0], spec[1])
        elif self._video_processing is not None:
            data = (spec[0], None)
        else:
            data = (None, spec[0])
        model = Seq2SeqModel(data_sequences=data, mode=mode, hparams=self._hparams, seed=self._seed)
        return Model(data=iterator, model=model, initializer=None, batch_size=batch_size)

    def _say(self, msg):
        if self._verbose and parallel.rank() == 0:
            print(msg)

    # ---- training (avsr.py:227-320) ---------------------------------------------------------------------------
    def train(self, logfile, num_epochs=400, try_restore_latest_checkpoint=False):
        checkpoint_dir = path.join(self._workdir, 'checkpoints', path.split(logfile)[-1])
        checkpoint_path = path.join(checkpoint_dir, 'checkpoint.ckp')
        makedirs(checkpoint_dir, exist_ok=True)
        makedirs(path.dirname(path.abspath(logfile)), exist_ok=True)
        tm = self._train_model
        last_epoch = 0
        if try_restore_latest_checkpoint is True:
            # avsr.py:241-249 swallows every restore error and trains from scratch; here only "there is no checkpoint"
            # does that - a checkpoint that exists but cannot be loaded is an error, not a silent restart at epoch 0
            latest_ckp = None
            try:
                latest_ckp = latest_checkpoint(checkpoint_dir)
            except Exception:
                latest_ckp = None
            if latest_ckp is None:
                self._say('Could not restore from checkpoint, training from scratch!\n')
            else:
                last_epoch = int(latest_ckp.split('-')[-1])
                tm.model.saver.restore(sess=None, save_path=latest_ckp)
                self._say('Restoring checkpoint from epoch {}\n'.format(last_epoch))
        self.last_error_rate = None
        with open(logfile if parallel.rank() == 0 else devnull, 'a') as f:
            for current_epoch in range(1, num_epochs):
                epoch = last_epoch + current_epoch
                tm.data.iterator_initializer()
                sum_loss, batches = 0.0, 0
                start = time.time()
                try:
                    while True:
                        tm.data.next()
                        batch_loss, global_norm = tm.model.train_step(tm.data.data_sequences())
                        sum_loss += batch_loss
                        self._say('batch: {}, batch loss: {:.2f}, gradient norm: {:.2f}'.format(
                            batches, batch_loss, global_norm))
                        batches += 1
                except OutOfRangeError:
                    pass
                self._say('epoch time: {}'.format(time.time() - start))
                f.write('Average batch_loss as epoch {} is {}\n'.format(epoch, sum_loss / max(batches, 1)))
                f.flush()
                if epoch % 10 == 0:
                    if parallel.rank() == 0:  # replicas hold identical parameters: rank 0 saves and evaluates
                        save_path = tm.model.saver.save(sess=None, save_path=checkpoint_path, global_step=epoch)
                        if self._evaluate_model is not None:
                            error_rate = self.evaluate(save_path, epoch)
                            for (k, v) in error_rate.items():
                                f.write(k + ': {:.4f}% '.format(v * 100))
                            f.write('\n')
                            f.flush()
                            self.last_error_rate = error_rate
                    parallel.barrier()

    def _write_alignment_images(self, model, names, outdir):
        """avsr.py:409-436: <file>.png (decoder attention; `_video` / `_audio` for the bimodal decoder) and <file>_av.png
        (cross-modal attention of the AV-Align encoder); pixel = 1 - alignment, rows = memory steps."""
        from .utils import write_png_gray
        dec = model._decoder.attention_summary
        arch = self._hparams.architecture
        for idx in range(len(names)):
            fname = path.join(outdir, names[idx].decode('utf-8'))
            makedirs(path.dirname(fname) or '.', exist_ok=True)
            if arch == 'bimodal':
                write_png_gray(fname + '_video.png', dec[0][idx, :, :, 0])
                write_png_gray(fname + '_audio.png', dec[1][idx, :, :, 0])
            else:
                write_png_gray(fname + '.png', dec[idx, :, :, 0])
                if arch == 'av_align':
                    write_png_gray(fname + '_av.png', model._audio_encoder.attention_summary[idx, :, :, 0])

    # ---- evaluation (avsr.py:322-512) ---------------------------------------------------------------------------
    def evaluate(self, checkpoint_path, epoch=None, alignments_outdir='./alignments/tmp/',
                 beam_graphs_outdir='./beam_graphs/tmp/'):
        em = self._evaluate_model
        em.model.saver.restore(sess=None, save_path=checkpoint_path)
        em.data.iterator_initializer()
        predictions_dict, labels_dict = {}, {}
        while True:
            try:
                em.data.next()
            except OutOfRangeError:
                break
            predicted = em.model.predict(em.data.data_sequences())
            names = em.data.inputs_filenames
            names = names[0] if isinstance(names, tuple) else names
            labels = em.data.labels.numpy()
            if self._write_attention_alignment is True:
                self._write_alignment_images(em.model, names, alignments_outdir)
            for idx in range(len(names)):
                file = names[idx].decode('utf-8')
                predictions_dict[file] = [self._unit_dict[int(sym)] for sym in predicted[idx]]
                labels_dict[file] = [self._unit_dict[int(sym)] for sym in labels[idx]]
        uer, uer_dict = compute_wer(predictions_dict, labels_dict)
        error_rate = {self._unit: uer}
        if self._unit == 'character':
            wer, _ = compute_wer(predictions_dict, labels_dict, split_words=True)
            error_rate['word'] = wer
        outdir = path.join(self._workdir, 'predictions', path.split(path.split(checkpoint_path)[0])[-1])
        makedirs(outdir, exist_ok=True)
        write_sequences_to_labelfile(predictions_dict, path.join(outdir, 'predicted_epoch_{}.mlf'.format(epoch)),
                                     labels_dict, uer_dict, sep=' ' if self._unit == 'phoneme' else '')
        return error_rate
