"""Data parallelism on real GPUs (NCCL): two ranks with B utterances each reproduce one rank with the 2B-utterance
batch - loss, global gradient norm, input batch-norm moving statistics and the parameters after the step
(reference semantics of one large batch: encoder.py:44-50 BN over the whole batch, seq2seq.py:165-171 loss
denominator = all target tokens, seq2seq.py:246 clip by the GLOBAL norm).  Needs two visible GPUs (`gpurun --gpus 2`);
skipped on a one-GPU box."""
import os
import socket

import numpy as np
import pytest
import torch

from tests.helpers import config_hparams, synthetic_batch, to_data_sequences, to_image_sequences

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _slice(batch, lo, hi):
    return {k: v[lo:hi] for k, v in batch.items()}


def _batch(hp, n):
    batch = synthetic_batch(hp, B=n, Ta=40, Tv=12, L=8, ragged=True)
    if 'audio_len' in batch:
        batch['audio_len'][:] = np.maximum(batch['audio_len'], 1)
    if hp.video_processing == 'resnet_cnn':  # lip crops: the front-end's fused batch norms all-reduce their statistics
        to_image_sequences(batch, hw=12)
    return batch


def _one_step(hp, ds, graph):
    from avsr_tf1_b200.seq2seq import Seq2SeqModel
    model = Seq2SeqModel(ds, 'train', hp, seed=2001)
    model.use_cuda_graph = graph
    out = [model.train_step(ds) for _ in range(2)]
    st = model.store
    return out, st.to_numpy('p'), {k: v.cpu().numpy() for k, v in st.state.items()}


def _worker(rank, world, port, cfg, over, B, graph, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        hp = config_hparams(cfg, **over)
        batch = _batch(hp, world * B)
        ds = to_data_sequences(_slice(batch, rank * B, (rank + 1) * B))
        out, params, state = _one_step(hp, ds, graph)
        if rank == 0:
            ret.put((out, params, state))
            ret.close()
            ret.join_thread()  # flushed before the hard exit below
        dist.barrier()
        torch.cuda.synchronize()
    except BaseException:
        import traceback
        traceback.print_exc()
        os._exit(1)
    # a captured CUDA graph holds NCCL work: tearing the process group down under it can hang until the NCCL watchdog
    # fires (10 minutes), so the worker leaves hard once every rank has reached the barrier (bench.py does the same)
    os._exit(0)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs (gpurun --gpus 2)')
@pytest.mark.parametrize('cfg,graph,over', [(5, False, {}), (5, True, {}), (2, False, {}), (4, False, {}),
                                            (3, False, dict(video_processing='resnet_cnn'))])
def test_two_ranks_equal_one_rank_with_the_whole_batch(cfg, graph, over):
    import torch.multiprocessing as mp
    B = 6
    ctx = mp.get_context('spawn')
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, cfg, over, B, graph, ret), daemon=True) for r in range(2)]
    for p in procs:
        p.start()
    try:
        import queue as _q
        got = None
        for _ in range(300):  # a worker that died (exit code != 0) ends the wait at once
            try:
                got = ret.get(timeout=1)
                break
            except _q.Empty:
                assert all(p.exitcode in (None, 0) for p in procs), [p.exitcode for p in procs]
        assert got is not None, 'no result from rank 0 within 300 s'
        out2, params2, state2 = got
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0, p.exitcode
    finally:
        for p in procs:  # never leave a worker behind: pytest would wait for it at exit
            if p.is_alive():
                p.kill()
    hp = config_hparams(cfg, **over)
    out1, params1, state1 = _one_step(hp, to_data_sequences(_batch(hp, 2 * B)), False)
    for (l2, g2), (l1, g1) in zip(out2, out1):
        assert abs(l2 - l1) <= 2e-4 * abs(l1), (l2, l1)          # same loss of the same global batch
        assert abs(g2 - g1) <= 2e-3 * g1, (g2, g1)                # same global norm (clip acts on it)
    # batch-norm moving statistics.  Behind the front-end the features reach the input normalisation through 13 tf32
    # convolutions whose fp32 statistics are summed by atomics in a different order on 2 x 6 and on 12 utterances
    cnn = over.get('video_processing') == 'resnet_cnn'
    for k, v in state1.items():
        assert np.allclose(state2[k], v, rtol=1e-3 if cnn else 1e-4, atol=2e-5 if cnn else 1e-6), \
            (k, float(np.abs(state2[k] - v).max()), float(np.abs(v).max()))
    lr_sum = sum(hp.learning_rate * (s + 1) / 750.0 for s in range(2))
    for k, v in params1.items():                                  # Adam's first steps are sign-like: see test_gpu_model
        if k.endswith(('moving_mean', 'moving_variance')):        # not trained: the moving statistics of the front-end
            assert np.allclose(params2[k], v, rtol=1e-3 if cnn else 1e-4, atol=2e-5 if cnn else 1e-6), \
                (k, float(np.abs(params2[k] - v).max()), float(np.abs(v).max()))
            continue
        diff = np.abs(params2[k] - v)
        assert diff.max() <= 2.2 * lr_sum + 1e-6 * np.abs(v).max() + 1e-7, (k, diff.max())
        # (a sign-like first Adam step flips where a gradient entry is rounding noise: a few percent of a large tensor, one
        # or two entries of a 16-entry batch-norm gamma)
        assert (diff > 0.05 * lr_sum + 1e-6 * np.abs(v).max()).mean() < max(3e-2, 2.5 / diff.size), k
