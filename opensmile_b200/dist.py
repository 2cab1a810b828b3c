"""Multi-GPU plumbing: the LLD path shards over independent utterances, one rank
per GPU, NO data-path collective.  torch.distributed only carries the counter / timing
reduction and (optionally) the gather of per-rank row counts so that rank 0 can lay out a
global row index.  Works with the `nccl` backend on GPUs and with `gloo` on CPUs (tests).
"""
import numpy as np


def shard_utterances(lengths, world_size, rank):
    """Greedy size-balanced static partition of utterances over ranks (longest first).

    lengths: per-utterance sample counts.  Returns the sorted indices owned by `rank`.  Every
    rank computes the same partition from the same lengths; no communication is needed."""
    lengths = np.asarray(lengths, dtype=np.int64)
    order = np.argsort(-lengths, kind="stable")
    load = np.zeros(world_size, dtype=np.int64)
    owner = np.empty(len(lengths), dtype=np.int64)
    for i in order:
        r = int(np.argmin(load))
        owner[i] = r
        load[r] += lengths[i]
    return np.nonzero(owner == rank)[0]


def reduce_counters(frames, seconds, dist=None, device=None):
    """(total frames over ranks, max seconds over ranks).  `dist` = torch.distributed or None."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return int(frames), float(seconds)
    import torch
    t = torch.tensor([float(frames)], dtype=torch.float64, device=device)
    m = torch.tensor([float(seconds)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return int(round(float(t[0]))), float(m[0])


def gather_row_counts(local_rows, dist=None, device=None):
    """Per-rank output row counts on every rank (all_gather of one int64)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [int(local_rows)]
    import torch
    mine = torch.tensor([int(local_rows)], dtype=torch.int64, device=device)
    out = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    return [int(x[0]) for x in out]
