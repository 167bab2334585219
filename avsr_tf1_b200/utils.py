"""Host-side metric (avsr/utils.py:4-46): integer Levenshtein distance and the
mean normalised error rate.  Results must be bit-exact with the reference file
(tests/test_metric_golden.py)."""
from __future__ import annotations


def _strip_extra_chars(prediction):
    return [value for value in prediction if value not in ('EOS', 'END', 'MASK')]


def levenshtein(ground_truth, prediction):
    """Two-row dynamic programme, O(min(n, m)) space."""
    a, b = ground_truth, prediction
    if len(a) > len(b):
        a, b = b, a
    row = list(range(len(a) + 1))
    for i, tok in enumerate(b, start=1):
        new = [i] + [0] * len(a)
        for j in range(1, len(a) + 1):
            cost = row[j - 1] if a[j - 1] == tok else row[j - 1] + 1
            new[j] = min(row[j] + 1, new[j - 1] + 1, cost)
        row = new
    return row[len(a)]


def compute_wer(predictions_dict, ground_truth_dict, split_words=False):
    total = 0
    err_dict = {}
    for fname, prediction in predictions_dict.items():
        prediction = _strip_extra_chars(prediction)
        ground_truth = _strip_extra_chars(ground_truth_dict[fname])
        if split_words is True:
            prediction = ''.join(prediction).split()
            ground_truth = ''.join(ground_truth).split()
        er = levenshtein(ground_truth, prediction) / float(len(ground_truth))
        total += er
        err_dict[fname] = er
    return total / (float(len(predictions_dict)) or 1), err_dict


def ids_to_symbols(ids, unit_dict):
    """avsr.py:400-405."""
    return [unit_dict[int(i)] for i in ids]


def write_sequences_to_labelfile(sequence_dict, fname, original_dict, error_dict, sep=''):
    """avsr/utils.py:49-59 (.mlf dump)."""
    with open(fname, 'w') as f:
        for k, v in sequence_dict.items():
            label_str = sep.join(_strip_extra_chars(v))
            truth = sep.join(_strip_extra_chars(original_dict[k]))
            f.write(' '.join([k, label_str, '[{}] [{:.3f}]'.format(truth, error_dict[k])]) + '\n')


def write_png_gray(fname, image):
    """8-bit greyscale PNG of a [H, W] array in [0, 1] (what tf.summary.image encodes for the alignment images,
    avsr.py:409-436); zlib + the PNG chunk format only."""
    import struct
    import zlib

    import numpy as np
    img = (np.clip(np.asarray(image, np.float64), 0.0, 1.0) * 255.0 + 0.5).astype(np.uint8)
    h, w = img.shape
    raw = b''.join(b'\x00' + img[r].tobytes() for r in range(h))

    def chunk(tag, data):
        return struct.pack('>I', len(data)) + tag + data + struct.pack('>I', zlib.crc32(tag + data) & 0xFFFFFFFF)
    with open(fname, 'wb') as f:
        f.write(b'\x89PNG\r\n\x1a\n' + chunk(b'IHDR', struct.pack('>IIBBBBB', w, h, 8, 0, 0, 0, 0)) +
                chunk(b'IDAT', zlib.compress(raw, 6)) + chunk(b'IEND', b''))
