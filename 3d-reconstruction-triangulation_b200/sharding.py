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


# ---- the classifier over a frame-sharded sequence ---------------------------------
# Candidate generation (fillCombinationQueue, DroneClassifier.cpp:156-198) depends on the frame alone, so every
# rank enumerates its contiguous frame range at once; linking (classifyDrones' path state, :119-135) is sequential,
# so the shards are linked in rank order and only the tracking state -- n_drones x (last point + 3-point tail),
# ~1.2 KB -- travels along the chain.

def slice_csr(det_offsets, dets_xy, n_cams: int, n_frames: int, f0: int, f1: int):
    """The CSR detections ([cam][frame+1] offsets into [n][2] pixel pairs) restricted to frames [f0, f1)."""
    import numpy as np
    o = np.asarray(det_offsets, np.int32).reshape(n_cams, n_frames + 1)
    xy = np.asarray(dets_xy, np.float64).reshape(-1, 2)
    offs = np.zeros((n_cams, f1 - f0 + 1), np.int32)
    parts, base = [], 0
    for c in range(n_cams):
        a, b = int(o[c, f0]), int(o[c, f1])
        offs[c] = o[c, f0:f1 + 1] - a + base
        parts.append(xy[a:b])
        base += b - a
    return offs.reshape(-1), (np.concatenate(parts) if parts else np.zeros((0, 2)))


def classify_chain(engines, mode, n_drones, det_offsets, dets_xy, n_frames, flags=0):
    """One process, one engine per shard (different GPUs, or the same one): enumerate every shard, then link them
    in order handing the state along.  Returns the dict of Engine.classify for the whole sequence."""
    import numpy as np
    world = len(engines)
    n_cams = len(engines[0].cameras)
    ranges = [(n_frames * g // world, n_frames * (g + 1) // world) for g in range(world)]
    for eng, (f0, f1) in zip(engines, ranges):
        offs, xy = slice_csr(det_offsets, dets_xy, n_cams, n_frames, f0, f1)
        eng.classify_begin(mode, n_drones, offs, xy, f1 - f0, flags)
    state, outs = None, []
    for eng in engines:
        r = eng.classify_finish(state)
        state = r["state"]
        outs.append(r)
    stats = {k: (max if k == "max_frontier" else sum)(o["stats"][k] for o in outs) for k in outs[0]["stats"]}
    return dict(paths=np.concatenate([o["paths"] for o in outs], axis=1), assign=np.concatenate([o["assign"] for o in outs], axis=1),
                phase=np.concatenate([o["phase"] for o in outs], axis=1), stats=stats)


def classify_sharded(engine, mode, n_drones, det_offsets, dets_xy, n_frames, rank: int, world: int, flags=0, group=None):
    """One process per GPU (torch.distributed): every rank enumerates its frame range, then rank g waits for rank
    g-1's tracking state (one point-to-point message), links its shard and passes the state on.  Returns this rank's
    shard of the outputs and its frame range; gather them with `gather_classified`."""
    import numpy as np
    import torch
    import torch.distributed as dist
    n_cams = len(engine.cameras)
    f0, f1 = n_frames * rank // world, n_frames * (rank + 1) // world
    offs, xy = slice_csr(det_offsets, dets_xy, n_cams, n_frames, f0, f1)
    engine.classify_begin(mode, n_drones, offs, xy, f1 - f0, flags)
    dev = torch.device("cuda", engine.device) if world > 1 and dist.get_backend(group) == "nccl" else torch.device("cpu")
    from . import lib
    nb = lib().tri_classify_state_bytes()
    state = None
    if rank > 0:
        buf = torch.empty(nb, dtype=torch.uint8, device=dev)
        dist.recv(buf, src=rank - 1, group=group)
        state = buf.cpu().numpy().tobytes()
    r = engine.classify_finish(state)
    if rank + 1 < world:
        dist.send(torch.frombuffer(bytearray(r["state"]), dtype=torch.uint8).to(dev), dst=rank + 1, group=group)
    r["frames"] = (f0, f1)
    return r


def gather_classified(r, n_frames: int, rank: int, world: int, group=None):
    """All ranks' shards of classify_sharded put together on every rank (paths / assign / phase in frame order)."""
    import numpy as np
    import torch.distributed as dist
    if world == 1:
        return r
    parts = [None] * world
    dist.all_gather_object(parts, {k: r[k] for k in ("paths", "assign", "phase", "stats", "frames")}, group=group)
    parts.sort(key=lambda q: q["frames"][0])
    stats = {k: (max if k == "max_frontier" else sum)(q["stats"][k] for q in parts) for k in parts[0]["stats"]}
    return dict(paths=np.concatenate([q["paths"] for q in parts], axis=1), assign=np.concatenate([q["assign"] for q in parts], axis=1),
                phase=np.concatenate([q["phase"] for q in parts], axis=1), stats=stats)
