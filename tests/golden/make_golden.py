"""Generates tests/golden/metric_vectors.json by importing the REFERENCE's own avsr/utils.py
(pure Python, no TensorFlow import) from /root/reference.  Run in the build container only; the
fixture travels, the reference does not.

    python tests/golden/make_golden.py
"""
import importlib.util
import json
import os
import random

REF = '/root/reference/avsr/utils.py'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'metric_vectors.json')

spec = importlib.util.spec_from_file_location('ref_utils', REF)
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

rng = random.Random(1234)
alphabet = [' ', "'"] + [chr(c) for c in range(ord('a'), ord('z') + 1)]


def rand_seq(n, alpha=alphabet):
    return [rng.choice(alpha) for _ in range(n)]


lev = []
# the known answers quoted in SURVEY.md 8c
for a, b in [('kitten', 'sitting'), ('abc', ''), ('', ''), ('flaw', 'lawn'), ('intention', 'execution')]:
    lev.append({'a': list(a), 'b': list(b), 'd': ref.levenshtein(list(a), list(b))})
for _ in range(300):  # random pairs incl. empty and very unequal lengths, small alphabets force collisions
    alpha = alphabet if rng.random() < 0.5 else alphabet[:3]
    a, b = rand_seq(rng.randint(0, 45), alpha), rand_seq(rng.randint(0, 45), alpha)
    if rng.random() < 0.3:  # near-identical strings
        b = list(a)
        for _ in range(rng.randint(0, 4)):
            if b and rng.random() < 0.5:
                del b[rng.randrange(len(b))]
            else:
                b.insert(rng.randint(0, len(b)), rng.choice(alpha))
    lev.append({'a': a, 'b': b, 'd': ref.levenshtein(a, b)})

wer = []
pred = {'a': list('the cat sat') + ['EOS', 'EOS'], 'b': list('hello wrld') + ['EOS', 'MASK']}
truth = {'a': list('the cat sat') + ['EOS'], 'b': list('hello world') + ['EOS']}
for split in (False, True):
    v, d = ref.compute_wer(pred, truth, split_words=split)
    wer.append({'pred': pred, 'truth': truth, 'split_words': split, 'value': v, 'per_file': d})
for case in range(40):
    n = rng.randint(1, 6)
    P, T = {}, {}
    for i in range(n):
        words = [''.join(rng.choice(alphabet[2:]) for _ in range(rng.randint(1, 6))) for _ in range(rng.randint(1, 7))]
        t = list(' '.join(words))
        p = list(t)
        for _ in range(rng.randint(0, 6)):
            if p and rng.random() < 0.5:
                del p[rng.randrange(len(p))]
            else:
                p.insert(rng.randint(0, len(p)), rng.choice(alphabet))
        P[f'utt{i}'] = p + ['EOS'] + ['MASK'] * rng.randint(0, 3)
        T[f'utt{i}'] = t + ['EOS']
    for split in (False, True):
        if split and any(not ''.join(ref._strip_extra_chars(v)).split() for v in T.values()):
            continue
        v, d = ref.compute_wer(P, T, split_words=split)
        wer.append({'pred': P, 'truth': T, 'split_words': split, 'value': v, 'per_file': d})

# io_utils.create_unit_dict needs TF only for other functions; restate-check via the character list
with open('/root/reference/avsr/misc/character_list') as f:
    units = f.read().splitlines()
unit_dict = {'MASK': 0, 'END': -1}
for idx, u in enumerate(units):
    unit_dict[u] = idx + 1
unit_dict['EOS'] = idx + 2
unit_dict['GO'] = idx + 3

json.dump({'source': 'georgesterpu/avsr-tf1 avsr/utils.py (levenshtein :22-42, compute_wer :4-19)',
           'levenshtein': lev, 'compute_wer': wer,
           'unit_dict_character': {str(v): k for k, v in unit_dict.items()}}, open(OUT, 'w'))
print('wrote', OUT, len(lev), 'levenshtein cases,', len(wer), 'compute_wer cases')
