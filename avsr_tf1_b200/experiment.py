"""run_experiment - drop-in for reference avsr/experiment.py:5-136: the curriculum the run_*.py scripts drive
(optional warm-up on short sentences, then per noise condition two stages at two learning rates, every stage a fresh
AVSR object that resumes from the run's latest checkpoint)."""
from __future__ import annotations

from os import path

from .avsr import AVSR


def _stage(logfile, num_epochs, separator, **avsr_kwargs):
    experiment = AVSR(**avsr_kwargs)
    experiment.train(logfile=logfile, num_epochs=num_epochs, try_restore_latest_checkpoint=True)
    with open(logfile, 'a') as f:
        f.write(separator * '=' + '\n')
    return experiment


def run_experiment(video_train_record=None, video_test_record=None, labels_train_record=None, labels_test_record=None,
                   audio_train_records=None, audio_test_records=None, unit='character', unit_list_file=None,
                   iterations=None, learning_rates=None, logfile='tmp_experiment', warmup_epochs=0, warmup_max_len=50,
                   input_modality='audio', logdir='./logs', **kwargs):
    full_logfile = path.join(logdir, logfile)
    common = dict(unit=unit, unit_file=unit_list_file, video_train_record=video_train_record,
                  video_test_record=video_test_record, labels_train_record=labels_train_record,
                  labels_test_record=labels_test_record, **kwargs)
    video_only = input_modality == 'video'
    last = None
    if warmup_epochs >= 1:  # experiment.py:24-51
        from os import makedirs
        makedirs(logdir, exist_ok=True)
        with open(full_logfile, 'a') as f:
            f.write('Warm up on short sentences up to {} tokens for {} epochs \n'.format(warmup_max_len, warmup_epochs))
        last = _stage(full_logfile, warmup_epochs, 5, learning_rate=learning_rates[0][0],
                      max_sentence_length=warmup_max_len,
                      audio_train_record=None if video_only else audio_train_records[0],
                      audio_test_record=None if video_only else audio_test_records[0], **common)
    if video_only:  # experiment.py:53-90
        conditions = [(learning_rates[0], iterations[0], None, None)]
    else:  # experiment.py:92-136: one pass per audio condition (clean, 10 dB, 0 dB, -5 dB in run_audio.py)
        conditions = list(zip(learning_rates, iterations, audio_train_records, audio_test_records))
    for lr, iters, audio_train, audio_test in conditions:
        for stage in (0, 1):
            last = _stage(full_logfile, iters[stage] + 1, 5 if stage == 0 else 20, learning_rate=lr[stage],
                          audio_train_record=audio_train, audio_test_record=audio_test, **common)
    return last
