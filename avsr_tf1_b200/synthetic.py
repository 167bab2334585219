"""Synthetic TFRecords in the reference's schema: audio `inputs` [Ta, 80] ~ N(0,1) (seed
1001), video `inputs` [Tv, 36, 36, 3] ~ U(-1,1) (seed 1002, the (v - 128) / 128 range of dataset_writer.py:537),
labels uniform in 1..28 (seed 1003); ragged=True draws lengths in [T/2, T] / [L/2, L] (seed 1004).
Written with the native writer (include/avsr_io.h), ids `utt%06d`."""
from __future__ import annotations

import os

import numpy as np

from .tfrecord import RecordWriter


def write_synthetic_records(outdir, n=4096, Ta=300, Tv=75, Fa=80, hw=36, channels=3, L=40, ragged=False,
                            with_aus=False, video=True, audio=True, prefix='synthetic', seed=0, n_classes=28):
    """Returns dict(video=path | None, audio=path | None, labels=path)."""
    os.makedirs(outdir, exist_ok=True)
    ra, rv, rl, rr, ru = (np.random.default_rng(s + seed) for s in (1001, 1002, 1003, 1004, 1005))
    paths = dict(video=os.path.join(outdir, prefix + '_video.tfrecord') if video else None,
                 audio=os.path.join(outdir, prefix + '_audio.tfrecord') if audio else None,
                 labels=os.path.join(outdir, prefix + '_labels.tfrecord'))
    wv = RecordWriter(paths['video']) if video else None
    wa = RecordWriter(paths['audio']) if audio else None
    wl = RecordWriter(paths['labels'])
    try:
        for i in range(n):
            sid = 'utt%06d' % i
            tv = int(rr.integers(max(1, Tv // 2), Tv + 1)) if ragged else Tv
            ta = tv * (Ta // Tv) if (ragged and video and audio) else (
                int(rr.integers(max(1, Ta // 2), Ta + 1)) if ragged else Ta)
            nl = int(rr.integers(max(1, L // 2), L + 1)) if ragged else L
            if wv is not None:
                aus = ru.uniform(0.0, 3.5, (tv, 2)).astype(np.float32) if with_aus else None
                wv.write_video(sid, rv.uniform(-1, 1, (tv, hw, hw, channels)).astype(np.float32), aus)
            if wa is not None:
                wa.write_feature(sid, ra.standard_normal((ta, Fa)).astype(np.float32))
            wl.write_labels(sid, rl.integers(1, n_classes + 1, nl), unit='character')
    finally:
        for w in (wv, wa, wl):
            if w is not None:
                w.close()
    return paths
