"""world_size-2 gloo test (CPU) of the data-parallel exchanges: summed batch-norm statistics, the global
loss denominator and the gradient all-reduce reproduce the single-rank large-batch values (checked with
the oracle's arithmetic)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import avsr_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from avsr_tf1_b200 import parallel
    rng = np.random.default_rng(0)
    B, T, F, V = 6, 5, 4, 7
    x = rng.standard_normal((B, T, F))
    logits = rng.standard_normal((B, T, V))
    lens = np.array([5, 2, 4, 3, 5, 1])
    tgt = rng.integers(0, V, (B, T))
    lo, hi = parallel.shard_batch(B)
    # (1) input batch-norm statistics: all-reduce [sum, sumsq], count = rows * world
    xs = torch.from_numpy(x[lo:hi].reshape(-1, F))
    sums = torch.cat([xs.sum(0), (xs * xs).sum(0)])
    parallel.allreduce_sum_(sums)
    count = xs.shape[0] * parallel.world_size()
    mean = sums[:F] / count
    var = sums[F:] / count - mean * mean
    # (2) loss denominator and (3) summed gradients
    n_tok = parallel.global_token_count(float(lens[lo:hi].sum()))
    w = O.sequence_mask(lens[lo:hi], T, np.float64)
    z = logits[lo:hi]
    logp = z - np.log(np.exp(z).sum(-1, keepdims=True))
    d = np.exp(logp)
    np.put_along_axis(d, tgt[lo:hi, :, None], np.take_along_axis(d, tgt[lo:hi, :, None], 2) - 1.0, axis=2)
    g_local = torch.from_numpy((d * (w / (n_tok + 1e-12))[:, :, None]).sum(0))  # a "parameter gradient"
    parallel.allreduce_sum_(g_local)
    if rank == 0:
        out.put((mean.numpy(), var.numpy(), n_tok, g_local.numpy()))
    dist.destroy_process_group()


def test_two_rank_exchanges_match_single_large_batch():
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    mean, var, n_tok, g = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    B, T, F, V = 6, 5, 4, 7
    x = rng.standard_normal((B, T, F))
    logits = rng.standard_normal((B, T, V))
    lens = np.array([5, 2, 4, 3, 5, 1])
    tgt = rng.integers(0, V, (B, T))
    _, _, m_ref, v_ref = O.batchnorm_train_fwd(x, np.ones(F), np.zeros(F))
    assert np.allclose(mean, m_ref) and np.allclose(var, v_ref)
    assert n_tok == float(lens.sum())
    _, d_ref = O.sequence_loss_fwd_bwd(logits, tgt, lens)
    assert np.allclose(g, d_ref.sum(0))


def test_shard_batch_covers_everything_once():
    from avsr_tf1_b200 import parallel
    for n in (1, 7, 8, 256, 1000):
        for w in (1, 2, 3, 8):
            spans = [parallel.shard_batch(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _shard_worker(rank, world, port, paths, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from avsr_tf1_b200 import io_utils, parallel
    from avsr_tf1_b200.hparams import create_unit_dict
    it = io_utils.make_iterator_from_two_records(
        paths['video'], paths['audio'], paths['labels'], batch_size=6, unit_dict=create_unit_dict(None), shuffle=True,
        bucket_width=3, seed=5, shuffle_buffer=8, prefetch=1, pin_memory=False,
        shard=(parallel.rank(), parallel.world_size()))
    steps = []
    for b in it:
        names = [n.decode() for n in b.labels_filenames.tolist()]
        # what a training step does between ranks: one collective per step must pair up on every rank
        t = torch.tensor([float(len(names))])
        parallel.allreduce_sum_(t)
        # input batch-norm statistics the way layers.BatchNormInput forms them under data parallelism: all-reduced
        # [sum, sumsq] over this rank's rows (padding included), row count = padded length x GLOBAL batch size
        audio = b.data_sequences()[1]
        x = torch.as_tensor(audio.inputs, dtype=torch.float64)
        n, t_pad, F = x.shape
        sums = torch.cat([x.reshape(-1, F).sum(0), (x.reshape(-1, F) ** 2).sum(0)])
        parallel.allreduce_sum_(sums)
        count = t_pad * audio.payload['global_batch_size']
        mean = (sums[:F] / count).numpy()
        var = (sums[F:] / count).numpy() - mean * mean
        steps.append((names, int(t.item()), t_pad, audio.payload['global_batch_size'], mean, var))
    parallel.barrier()
    out.put((rank, steps))
    dist.destroy_process_group()


def test_record_iterator_shards_pair_up_across_ranks(tmp_path):
    """Two gloo ranks walk the same shuffled, bucketed epoch of synthetic TFRecords: same number of steps (their per-step
    collective pairs up), disjoint slices, and together every utterance of every kept global batch exactly once."""
    from avsr_tf1_b200.synthetic import write_synthetic_records
    paths = write_synthetic_records(str(tmp_path), n=31, Ta=16, Tv=8, Fa=4, hw=3, L=4, ragged=True)
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, paths, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(out.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    s0, s1 = got[0], got[1]
    assert len(s0) == len(s1) > 0
    # the single-rank iterator over the same epoch: its batches are the global batches the two ranks share
    from avsr_tf1_b200 import io_utils
    from avsr_tf1_b200.hparams import create_unit_dict
    whole = io_utils.make_iterator_from_two_records(
        paths['video'], paths['audio'], paths['labels'], batch_size=6, unit_dict=create_unit_dict(None), shuffle=True,
        bucket_width=3, seed=5, shuffle_buffer=8, prefetch=0, pin_memory=False)
    ref = {}
    for b in whole:
        x = np.asarray(b.data_sequences()[1].inputs, np.float64)
        key = frozenset(n.decode() for n in b.labels_filenames.tolist())
        ref[key] = (x.shape[1], x.shape[0], x.reshape(-1, x.shape[2]).mean(0), x.reshape(-1, x.shape[2]).var(0))
    seen, uneven = [], 0
    for (n0, tot0, tp0, gb0, m0, v0), (n1, tot1, tp1, gb1, m1, v1) in zip(s0, s1):
        assert tot0 == tot1 == len(n0) + len(n1)      # the all-reduce saw both slices of the same global batch
        assert not set(n0) & set(n1) and abs(len(n0) - len(n1)) <= 1
        uneven += len(n0) != len(n1)
        # both ranks padded to the global batch's longest sequence and know its size: the all-reduced statistics are
        # those of the one large batch, also when the batch does not divide evenly (count was T*B*world)
        t_ref, b_ref, m_ref, v_ref = ref[frozenset(n0 + n1)]
        assert tp0 == tp1 == t_ref and gb0 == gb1 == b_ref == len(n0) + len(n1)
        assert np.allclose(m0, m_ref) and np.allclose(m1, m_ref) and np.allclose(v0, v_ref) and np.allclose(v1, v_ref)
        seen += n0 + n1
    assert uneven > 0  # the epoch contains batches that do not divide evenly between the ranks
    assert len(seen) == len(set(seen))
    assert len(seen) >= 31 - 2 * 4  # only batches smaller than the world (at most one per bucket) are dropped
