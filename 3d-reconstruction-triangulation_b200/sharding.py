"""Frame-range sharding of the batch path over the GPUs of one box.

Frames are independent (MatrixTriangulator.cpp:80-97), so rank g of G owns the contiguous range
[g*N/G, (g+1)*N/G) of the global frame index space, runs the same kernels on it and nothing is
exchanged on the data path.  The only collective is the optional final gather of the float3 points
(`gather_points`: one all_gather over NCCL/NVLink, or gloo in the CPU tests)."""
from __future__ import annotations


def shard_range(n_frames: int, rank: int, world: int):
    """[begin, end) of rank's frames: contiguous, ordered by rank, sizes differ by at most one frame pair
    (boundaries are kept even so every shard's rows stay 16-byte aligned for the vector loads)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")

    def cut(g):
        b = (n_frames * g) // world
        return min(n_frames, b + (b & 1)) if g < world else n_frames

    return cut(rank), cut(rank + 1)


def shard_sizes(n_frames: int, world: int):
    return [e - b for b, e in (shard_range(n_frames, g, world) for g in range(world))]


def gather_points(local, n_frames: int, rank: int, world: int, group=None):
    """All ranks get the [n_frames, 3] points in global frame order.  `local` is this rank's
    [shard, 3] tensor (any device the process group supports)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local
    sizes = shard_sizes(n_frames, world)
    assert local.shape[0] == sizes[rank], (local.shape, sizes, rank)
    m = max(sizes)
    if all(s == m for s in sizes):  # the common case: one all_gather straight into the result
        out = torch.empty((n_frames,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:sizes[rank]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)


def first_bad_frame(local_bad: int, begin: int, device, world: int, group=None):
    """Global index of the first frame with < 2 views over all shards (-1 if none): a MIN all-reduce."""
    import torch
    import torch.distributed as dist
    big = 1 << 62
    t = torch.tensor([big if local_bad < 0 else begin + local_bad], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    v = int(t.item())
    return -1 if v == big else v


def open_root_buffer(engine, nbytes: int, rank: int, world: int, root: int = 0, group=None):
    """`root` allocates `nbytes` of device memory through the engine; every other rank maps it (CUDA IPC, peer
    access enabled lazily), so kernels on any rank can store their shard of the result directly into the root's
    array over NVLink -- the gather fused into the producing kernel instead of a collective after it.
    Returns the raw device address valid in THIS process."""
    import torch.distributed as dist
    box = [None]
    ptr = None
    if rank == root:
        ptr = engine.device_alloc(nbytes)
        box = [engine.ipc_export(ptr)]
    dist.broadcast_object_list(box, src=root, group=group)
    if rank != root:
        ptr = engine.ipc_open(box[0])
    return ptr


def close_root_buffer(engine, ptr, rank: int, root: int = 0):
    if rank == root:
        engine.device_free(ptr)
    else:
        engine.ipc_close(ptr)
