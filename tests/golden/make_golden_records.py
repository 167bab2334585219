"""Generates tests/golden/sequence_examples.{tfrecord,json}: TFRecord files of tf.train.SequenceExample protos in the
schema of the reference's dataset_writer.py (:290-311 labels, :439-458 features, :461-498 video + aus), serialised by
google.protobuf (NOT by this repo's writer) and framed with tensorboard's masked crc32c, together with the decoded
content as JSON.  The native reader must reproduce the JSON from the bytes (tests/test_tfrecord_io.py), whether or
not protobuf / tensorboard are installed where the tests run.

    python tests/golden/make_golden_records.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.test_tfrecord_io import _example_classes, _frame  # noqa: E402  (protobuf messages, tensorboard framing)

SequenceExample = _example_classes()
rng = np.random.default_rng(20261017)
content = {'feature': [], 'video': [], 'labels': []}
blobs = {'feature': b'', 'video': b'', 'labels': b''}
for i, (steps, n_labels) in enumerate([(3, 4), (1, 1), (6, 9)]):
    sid = ('s%02d/utt_%d' % (i, i)).encode()
    x = rng.standard_normal((steps * 4, 5)).astype(np.float32)
    ex = SequenceExample()
    ex.context.feature['input_length'].int64_list.value.append(len(x))
    ex.context.feature['input_size'].int64_list.value.append(5)
    ex.context.feature['filename'].bytes_list.value.append(sid)
    for row in x:
        ex.feature_lists.feature_list['inputs'].feature.add().float_list.value.extend(row.tolist())
    blobs['feature'] += _frame(ex.SerializeToString())
    content['feature'].append({'filename': sid.decode(), 'inputs': x.tolist()})

    frames = (rng.integers(0, 256, (steps, 3, 2, 3)).astype(np.float32) - 128) / 128  # [T, height 3, width 2, channels 3]
    aus = rng.uniform(0, 5, (steps, 2)).astype(np.float32)
    ex = SequenceExample()
    ex.context.feature['input_length'].int64_list.value.append(steps)
    ex.context.feature['width'].int64_list.value.append(2)
    ex.context.feature['height'].int64_list.value.append(3)
    ex.context.feature['channels'].int64_list.value.append(3)
    ex.context.feature['filename'].bytes_list.value.append(sid)
    for fr, au in zip(frames, aus):
        ex.feature_lists.feature_list['inputs'].feature.add().float_list.value.extend(fr.flatten().tolist())
        ex.feature_lists.feature_list['aus'].feature.add().float_list.value.extend(au.tolist())
    blobs['video'] += _frame(ex.SerializeToString())
    content['video'].append({'filename': sid.decode(), 'inputs': frames.reshape(steps, -1).tolist(), 'aus': aus.tolist()})

    y = rng.integers(1, 29, n_labels)
    ex = SequenceExample()
    ex.context.feature['unit'].bytes_list.value.append(b'character')
    ex.context.feature['labels_length'].int64_list.value.append(n_labels)
    ex.context.feature['filename'].bytes_list.value.append(sid)
    for v in y:
        ex.feature_lists.feature_list['labels'].feature.add().int64_list.value.append(int(v))
    blobs['labels'] += _frame(ex.SerializeToString())
    content['labels'].append({'filename': sid.decode(), 'labels': [int(v) for v in y]})

for k, b in blobs.items():
    open(os.path.join(HERE, 'sequence_examples_%s.tfrecord' % k), 'wb').write(b)
json.dump(content, open(os.path.join(HERE, 'sequence_examples.json'), 'w'))
print({k: len(b) for k, b in blobs.items()})
