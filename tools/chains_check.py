"""Diagnostic: config 2 (BiLSTM) training steps with the two stacks on one stream vs side by side, eager vs graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from avsr_tf1_b200.seq2seq import Seq2SeqModel
from tests.helpers import config_hparams, synthetic_batch, to_data_sequences

hp = config_hparams(2)
batches = [synthetic_batch(hp, B=4, Ta=30, Tv=10, L=6, ragged=True, seed=s) for s in range(3)]
for b in batches:
    b['labels_len'][0] = 7
res = {}
for serial in (True, False):
    for graph in (False, True):
        m = Seq2SeqModel(to_data_sequences(batches[0]), 'train', hp, seed=2001)
        m.use_cuda_graph, m.serial_chains = graph, serial
        out = [m.train_step(to_data_sequences(b)) for b in batches]
        res[(serial, graph)] = out
        print('serial' if serial else 'parallel', 'graph' if graph else 'eager', ['%.6f %.6f' % o for o in out], flush=True)
