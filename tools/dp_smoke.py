"""2-rank data-parallel smoke test with hang diagnostics (run under torchrun)."""
import faulthandler, os, sys, time
faulthandler.dump_traceback_later(70, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
rank, local, world = int(os.environ['RANK']), int(os.environ['LOCAL_RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
def log(*a):
    print(f'[rank {rank} {time.time() % 1000:.1f}]', *a, flush=True)
from avsr_tf1_b200.seq2seq import Seq2SeqModel
from tests.helpers import config_hparams, synthetic_batch, to_data_sequences
hp = config_hparams(5)
full = synthetic_batch(hp, B=8, Ta=40, Tv=12, L=8, ragged=True)
lo, hi = rank * 4, rank * 4 + 4
mine = {k: v[lo:hi] for k, v in full.items()}
# identical padded label width on every rank (graphs are keyed by shape, values differ)
log('building model')
for graph in (False, True):
    m = Seq2SeqModel(to_data_sequences(mine), 'train', hp, seed=2001)
    m.use_cuda_graph = graph
    for s in range(3):
        out = m.train_step(to_data_sequences(mine))
        torch.cuda.synchronize()
        log('graph' if graph else 'eager', 'step', s, 'loss %.6f gnorm %.6f' % out)
# the reference-default training graph under data parallelism: dropout + scheduled sampling + AU head
from tests.helpers import add_aus
hp2 = config_hparams(5, use_dropout=True, sampling_probability_outputs=0.1, regress_aus=True)
full2 = add_aus(dict(full))
mine2 = {k: v[lo:hi] for k, v in full2.items()}
for graph in (False, True):
    m = Seq2SeqModel(to_data_sequences(mine2), 'train', hp2, seed=2001)
    m.use_cuda_graph = graph
    for s in range(2):
        out = m.train_step(to_data_sequences(mine2))
        torch.cuda.synchronize()
        log('default-graph', 'graph' if graph else 'eager', 'step', s, 'loss %.6f gnorm %.6f au %.6f' % (out + (m.au_loss,)),
            'rng', m.rng_words())
        assert all(np.isfinite(out))
if rank == 0:
    # single-rank large batch reference: same loss / grad-norm at step 0
    dist.barrier()
else:
    dist.barrier()
log('done')
# captured graphs hold NCCL work: tearing the process group down under them can hang (see bench.py finish())
sys.stdout.flush()
torch.cuda.synchronize()
os._exit(0)
