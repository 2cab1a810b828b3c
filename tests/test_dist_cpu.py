"""world_size-2 gloo test of the multi-GPU host logic (no GPU): static utterance sharding is a
partition, every rank derives the same one without communication, the per-rank frame counts sum
to the single-rank count (description-only plans), and the counter / timing reductions work."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from opensmile_b200 import Plan, components_mfcc12_0_d_a
from opensmile_b200.dist import gather_functionals, gather_row_counts, reduce_counters, shard_utterances


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, lengths, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    plan = Plan(components_mfcc12_0_d_a(16000.0), "lld", device=-1)     # description-only: no GPU here
    mine = shard_utterances(lengths, world, rank)
    off = np.zeros(len(mine) + 1, np.int64)
    off[1:] = np.cumsum(np.asarray(lengths)[mine])
    rows = int(plan.frame_offsets(off)[-1])
    total, tmax = reduce_counters(rows, 0.5 + rank, dist)
    counts = gather_row_counts(rows, dist)
    # the functionals gather: row i of the global matrix is [i, 2 i, 3 i] whichever rank owns utterance i
    local = np.stack([np.array([i, 2 * i, 3 * i], np.float32) for i in mine]) if len(mine) else np.zeros((0, 3), np.float32)
    glob = gather_functionals(local, mine, len(lengths), dist)
    ok = None if glob is None else bool(np.array_equal(glob.numpy(), np.arange(len(lengths), dtype=np.float32)[:, None] * np.array([1, 2, 3], np.float32)))
    q.put((rank, mine.tolist(), rows, total, tmax, counts, ok))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_gloo():
    rng = np.random.default_rng(0)
    lengths = rng.integers(300, 90000, size=37).tolist()
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, lengths, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    owned = sorted(i for r in res for i in r[1])
    assert owned == list(range(len(lengths)))                       # a partition
    plan = Plan(components_mfcc12_0_d_a(16000.0), "lld", device=-1)
    expect = sum(plan.num_frames(n) for n in lengths)
    assert res[0][3] == res[1][3] == expect == res[0][2] + res[1][2]     # SUM reduction == single-rank count
    assert res[0][4] == res[1][4] == 1.5                             # MAX over ranks
    assert res[0][5] == res[1][5] == [res[0][2], res[1][2]]
    assert res[0][6] is True and res[1][6] is None                   # rank 0 holds the gathered functionals matrix, in global order
    loads = [sum(lengths[i] for i in r[1]) for r in res]
    assert abs(loads[0] - loads[1]) <= max(lengths)                  # balanced


def test_shard_is_deterministic_and_balanced():
    lengths = [80240] * 2000
    parts = [shard_utterances(lengths, 8, r) for r in range(8)]
    assert sorted(np.concatenate(parts).tolist()) == list(range(2000))
    assert all(len(p) == 250 for p in parts)
