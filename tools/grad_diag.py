"""Diagnostic: per-parameter gradient errors of one configuration against the oracle (tensor-core mode)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import avsr_oracle as O
from avsr_tf1_b200 import ops
from tests.helpers import cast_batch
from tests.test_gpu_model import build, oracle_for

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 5
att = sys.argv[2] if len(sys.argv) > 2 else 'bahdanau'
ops.set_tensor_cores(True)
hp, batch, ds, model = build(cfg, dict(attention_type=((att,), (att,))), B=4, Ta=40, Tv=12, L=8)
om, P = oracle_for(hp, model)
loss_ref, G_ref, rec = om.loss_and_grads(cast_batch(batch, np.float64))
model.feed(ds); model._set_step_scalars(); model.forward_backward(); model.finish_gradients()
loss, gnorm = model.fetch_scalars()
print('loss', loss, loss_ref, 'gnorm', gnorm, O.global_norm(G_ref))
G = model.store.to_numpy('g')
gmax = max(np.abs(g).max() for g in G_ref.values())
for name, g_ref in G_ref.items():
    scale = max(np.abs(g_ref).max(), 1e-3 * gmax)
    err = np.abs(G[name].astype(np.float64) - g_ref).max() / scale
    print('%-70s %.3e %s' % (name, err, '<<<' if err > 1.5e-2 else ''))
