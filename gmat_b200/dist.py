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


def bind_rank_to_cores(local_rank, local_world):
    """Give every rank of a node its own slice of the host cores BEFORE it allocates pinned memory, so that the
    staging buffers of the eight processes are first-touched (and their copy threads run) on distinct cores, inside
    the CPU set NVML reports as local to the rank's GPU when that is known.  Returns the core list (or None)."""
    try:
        allowed = sorted(os.sched_getaffinity(0))
    except AttributeError:
        return None
    near = None
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (max(allowed) // 64) + 1)
        near = [c for c in allowed if (words[c // 64] >> (c % 64)) & 1]
    except Exception:
        near = None
    pool = near if near else allowed
    per = max(1, len(pool) // max(local_world, 1))
    mine = pool[(local_rank * per) % len(pool):][:per] or pool
    try:
        os.sched_setaffinity(0, set(mine))
    except OSError:
        return None
    return mine


def pipelined_scatter_compute_gather(full_in, in_frame_bytes, out_frame_bytes, n_frames, compute, chunk=8, root=0, device=None):
    """SURVEY 8e report (2): the root rank holds the whole input batch; every rank receives its shard in chunks of
    `chunk` frames, transforms chunk k while chunk k+1 is still arriving, and sends each result chunk back as soon
    as it is ready.  Point-to-point only (NCCL send/recv over NVLink on GPUs, gloo on the CPU tests); `compute(src,
    dst, nframes)` works on uint8 tensors.  Returns the full output on the root, None elsewhere."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    a, b = shard_range(n_frames, rank, world)
    mine_n = b - a
    if world == 1:
        out = torch.empty(n_frames * out_frame_bytes, dtype=torch.uint8, device=full_in.device)
        for c0 in range(0, n_frames, chunk):
            n = min(chunk, n_frames - c0)
            compute(full_in[c0 * in_frame_bytes:(c0 + n) * in_frame_bytes], out[c0 * out_frame_bytes:(c0 + n) * out_frame_bytes], n)
        return out
    dev = device if device is not None else (full_in.device if full_in is not None else None)
    full_out = torch.empty(n_frames * out_frame_bytes, dtype=torch.uint8, device=dev) if rank == root else None
    my_in = torch.empty(mine_n * in_frame_bytes, dtype=torch.uint8, device=dev) if rank != root else full_in[a * in_frame_bytes:b * in_frame_bytes]
    my_out = torch.empty(mine_n * out_frame_bytes, dtype=torch.uint8, device=dev) if rank != root else full_out[a * out_frame_bytes:b * out_frame_bytes]
    nchunks = (max(shard_range(n_frames, r, world)[1] - shard_range(n_frames, r, world)[0] for r in range(world)) + chunk - 1) // chunk
    pending = []
    # inputs: chunk-major, so that every rank gets its first chunk before anybody gets a second one
    recv_in = []
    for k in range(nchunks):
        ops = []
        if rank == root:
            for r in range(world):
                if r == root:
                    continue
                ra, rb = shard_range(n_frames, r, world)
                f0, f1 = ra + k * chunk, min(ra + (k + 1) * chunk, rb)
                if f1 > f0:
                    ops.append(dist.P2POp(dist.isend, full_in[f0 * in_frame_bytes:f1 * in_frame_bytes], r))
        else:
            f0, f1 = k * chunk, min((k + 1) * chunk, mine_n)
            if f1 > f0:
                ops.append(dist.P2POp(dist.irecv, my_in[f0 * in_frame_bytes:f1 * in_frame_bytes], root))
        recv_in.append(dist.batch_isend_irecv(ops) if ops else [])
    # compute chunk k as soon as it is here; its result leaves while chunk k+1 is converted
    for k in range(nchunks):
        for w in recv_in[k]:
            w.wait()
        f0, f1 = k * chunk, min((k + 1) * chunk, mine_n)
        if f1 > f0:
            compute(my_in[f0 * in_frame_bytes:f1 * in_frame_bytes], my_out[f0 * out_frame_bytes:f1 * out_frame_bytes], f1 - f0)
        ops = []
        if rank == root:
            for r in range(world):
                if r == root:
                    continue
                ra, rb = shard_range(n_frames, r, world)
                g0, g1 = ra + k * chunk, min(ra + (k + 1) * chunk, rb)
                if g1 > g0:
                    ops.append(dist.P2POp(dist.irecv, full_out[g0 * out_frame_bytes:g1 * out_frame_bytes], r))
        elif f1 > f0:
            ops.append(dist.P2POp(dist.isend, my_out[f0 * out_frame_bytes:f1 * out_frame_bytes], root))
        if ops:
            pending += dist.batch_isend_irecv(ops)
    for w in pending:
        w.wait()
    return full_out
