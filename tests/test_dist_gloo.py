"""world_size-2 gloo test (CPU) of the frame-sharding harness: scatter -> per-rank work -> gather."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gmat_b200.dist import gather_frames, scatter_frames, shard_by_cost, shard_range


def test_shard_range_partitions_every_frame_once():
    for n in (0, 1, 7, 8, 256, 1000):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                a, b = shard_range(n, r, world)
                got += list(range(a, b))
                assert b - a in (n // world, n // world + 1)
            assert got == list(range(n))


def test_shard_by_cost_balances_mixed_sizes():
    costs = [1920 * 1080] * 40 + [3840 * 2160] * 40 + [7680 * 4320] * 40        # BASELINE C5 mix
    parts = shard_by_cost(costs, 8)
    assert sorted(sum(parts, [])) == list(range(120))
    loads = [sum(costs[i] for i in p) for p in parts]
    assert max(loads) / min(loads) < 1.1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_frames, fb, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = None
    if rank == 0:
        full = torch.arange(n_frames * fb, dtype=torch.int64).remainder(251).to(torch.uint8)
    mine = scatter_frames(full, fb, n_frames, root=0)
    a, b = shard_range(n_frames, rank, world)
    assert mine.numel() == (b - a) * fb
    out = 255 - mine                                   # the per-rank "transform" (independent per frame)
    res = gather_frames(out, fb, n_frames, root=0)
    if rank == 0:
        q.put(bool(torch.equal(res, 255 - full)))
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == world
    dist.barrier()
    dist.destroy_process_group()


def test_scatter_gather_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 7, 1000, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def _pipe_worker(rank, world, port, n_frames, fin, fout, chunk, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gmat_b200.dist import pipelined_scatter_compute_gather
    full = None
    if rank == 0:
        full = torch.arange(n_frames * fin, dtype=torch.int64).remainder(253).to(torch.uint8)
    calls = []

    def compute(src, dst, n):          # a per-frame transform that shrinks the frame (like 4K -> 1080p)
        calls.append(n)
        dst.copy_((src.view(n, fin)[:, :fout] ^ 0x5A).reshape(-1))
    res = pipelined_scatter_compute_gather(full, fin, fout, n_frames, compute, chunk=chunk, root=0)
    a, b = shard_range(n_frames, rank, world)
    assert sum(calls) == b - a and all(c <= chunk for c in calls)
    if rank == 0:
        q.put(bool(torch.equal(res, (full.view(n_frames, fin)[:, :fout] ^ 0x5A).reshape(-1))))
    dist.barrier()
    dist.destroy_process_group()


def test_pipelined_scatter_compute_gather_world2_gloo():
    """the chunked scatter -> transform -> gather pipeline of bench.py's NVLink end-to-end figure, on gloo"""
    ctx = mp.get_context("spawn")
    for n_frames, chunk in ((21, 4), (5, 8), (16, 8)):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_pipe_worker, args=(r, 2, port, n_frames, 600, 150, chunk, q)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        assert q.get(timeout=5) is True


def test_bind_rank_to_cores_partitions_the_allowed_set():
    from gmat_b200.dist import bind_rank_to_cores
    before = os.sched_getaffinity(0)
    try:
        got = []
        for r in range(4):
            os.sched_setaffinity(0, before)          # every rank is its own process in real life
            got.append(bind_rank_to_cores(r, 4))
        os.sched_setaffinity(0, before)
        got = [g for g in got if g]
        assert got, "no affinity support"
        assert all(set(g) <= before for g in got)
        if len(before) >= 4:
            assert len(set().union(*[set(g) for g in got])) == sum(len(g) for g in got)      # disjoint
    finally:
        os.sched_setaffinity(0, before)
