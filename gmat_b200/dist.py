"""Frame-sharded data parallelism: frames are independent units, so a batch is cut by
batch index across the ranks of one node (one process per GPU, torch.distributed); there is no
collective inside the transform.  NCCL (or gloo on CPU, for the tests) is used only to scatter the
input batch from a root rank and to gather the results."""
import os

import torch
import torch.distributed as dist


def init(backend=None):
    """initialise from the torchrun environment; returns (rank, world, local_rank)"""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(n_frames, rank, world):
    """contiguous, balanced shard of frame indices: the first n % world ranks take one extra frame"""
    base, extra = divmod(n_frames, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_by_cost(costs, world):
    """greedy longest-processing-time assignment for mixed frame sizes (BASELINE C5);
    returns a list of frame-index lists, one per rank"""
    order = sorted(range(len(costs)), key=lambda i: -costs[i])
    loads = [0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: loads[k])
        out[r].append(i); loads[r] += costs[i]
    return [sorted(x) for x in out]


def scatter_frames(full, frame_bytes, n_frames, root=0, device=None):
    """root holds `full` (uint8 tensor of n_frames * frame_bytes); every rank gets its shard"""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    a, b = shard_range(n_frames, rank, world)
    if world == 1:
        return full[a * frame_bytes:b * frame_bytes]
    mine = torch.empty((b - a) * frame_bytes, dtype=torch.uint8, device=device)
    ops = []
    if rank == root:
        for r in range(world):
            ra, rb = shard_range(n_frames, r, world)
            chunk = full[ra * frame_bytes:rb * frame_bytes]
            if r == root:
                mine.copy_(chunk)
            elif rb > ra:
                ops.append(dist.P2POp(dist.isend, chunk, r))
    elif b > a:
        ops.append(dist.P2POp(dist.irecv, mine, root))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return mine


def gather_frames(mine, frame_bytes, n_frames, root=0, device=None):
    """inverse of scatter_frames: root returns the full batch in frame order, others None"""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    if world == 1:
        return mine
    full = torch.empty(n_frames * frame_bytes, dtype=torch.uint8, device=device) if rank == root else None
    ops = []
    a, b = shard_range(n_frames, rank, world)
    if rank == root:
        for r in range(world):
            ra, rb = shard_range(n_frames, r, world)
            dst = full[ra * frame_bytes:rb * frame_bytes]
            if r == root:
                dst.copy_(mine)
            elif rb > ra:
                ops.append(dist.P2POp(dist.irecv, dst, r))
    elif b > a:
        ops.append(dist.P2POp(dist.isend, mine, root))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return full
