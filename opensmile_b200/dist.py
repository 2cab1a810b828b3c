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


def gather_functionals(local_rows, local_indices, n_total, dist=None, device=None, dst=0):
    """The one data exchange of a sharded extraction (BASELINE.json north_star: "NCCL only for the final functionals
    reduction"): every rank holds one functionals row per utterance it owns (`local_rows` [n_local, K] float32 tensor on
    `device`, `local_indices` = the global utterance indices of those rows); rank `dst` receives the [n_total, K] matrix in
    global utterance order, the other ranks get None.  The reference has no counterpart (it summarises one file per process
    and the sink appends to a shared ARFF / CSV file); this replaces that file-level merge by one gather over NVLink:
    all_gather of the row counts, then gather of the (padded) rows and of their indices (NCCL on GPUs, gloo in the CPU tests).
    """
    import torch
    local_rows = torch.as_tensor(local_rows, dtype=torch.float32, device=device)
    idx = torch.as_tensor(np.asarray(local_indices, dtype=np.int64), device=device)
    K = int(local_rows.shape[1]) if local_rows.dim() == 2 else 0
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        out = torch.zeros((n_total, K), dtype=torch.float32, device=device)
        out[idx] = local_rows
        return out
    world, rank = dist.get_world_size(), dist.get_rank()
    counts = gather_row_counts(int(local_rows.shape[0]), dist, device)
    cap = max(max(counts), 1)
    pad_rows = torch.zeros((cap, K), dtype=torch.float32, device=device)
    pad_idx = torch.full((cap,), -1, dtype=torch.int64, device=device)
    pad_rows[:local_rows.shape[0]] = local_rows
    pad_idx[:idx.shape[0]] = idx
    rows_list = [torch.zeros_like(pad_rows) for _ in range(world)] if rank == dst else None
    idx_list = [torch.zeros_like(pad_idx) for _ in range(world)] if rank == dst else None
    dist.gather(pad_rows, rows_list, dst=dst)
    dist.gather(pad_idx, idx_list, dst=dst)
    if rank != dst:
        return None
    out = torch.zeros((n_total, K), dtype=torch.float32, device=device)
    for r in range(world):
        n = counts[r]
        out[idx_list[r][:n]] = rows_list[r][:n]
    return out
