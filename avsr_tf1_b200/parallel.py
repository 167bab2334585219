"""Data-parallel plumbing (one process per GPU, torch.distributed; NCCL over NVLink on the GPU box,
gloo in the CPU tests).  Training shards by utterance; the only exchanges are
  * the flat gradient buffer (one all-reduce per step),
  * the [2F] input batch-norm sums of each stream, forward and backward (exact large-batch statistics,
    reference encoder.py:44-50), and
  * the token count of the loss denominator (seq2seq.sequence_loss, seq2seq.py:165-171),
so N ranks with per-rank batch B reproduce one rank with batch N*B."""
from __future__ import annotations

import torch
import torch.distributed as dist


def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def allreduce_sum_(t: torch.Tensor) -> torch.Tensor:
    """In-place sum over ranks (no-op on a single rank)."""
    if world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def barrier():
    if world_size() > 1:
        dist.barrier()


def global_token_count(local_tokens: float, device=None) -> float:
    """Sum of labels_len over all ranks: the loss denominator of the global batch."""
    if world_size() == 1:
        return float(local_tokens)
    t = torch.tensor([local_tokens], dtype=torch.float64, device=device)
    dist.all_reduce(t)
    return float(t.item())


def shard_batch(n_utterances: int, r: int = None, w: int = None):
    """Contiguous utterance range [lo, hi) of rank r (the reference has no sharding: single device)."""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    per, rem = divmod(n_utterances, w)
    lo = r * per + min(r, rem)
    return lo, lo + per + (1 if r < rem else 0)
