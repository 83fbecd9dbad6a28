"""world_size-2 gloo worker for tests/test_sharding_cpu.py (run under torch.distributed.run)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tri_b200  # noqa: E402,F401
from tri_b200 import sharding as SH  # noqa: E402
from tri_b200 import synthetic as S  # noqa: E402

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
cams = S.ring_rig(8)
for n in (1000, 1001, 7, 2, 4096):
    b, e = SH.shard_range(n, rank, world)
    # each rank generates ITS frames from the global index space and "solves" them with a stand-in that is
    # a pure function of the frame's pixels (the CUDA solve is covered by the -m gpu tests)
    xy = S.generate_frames(cams, e - b, frame0=b)
    local = torch.stack([xy[:, :, 0].sum(0), xy[:, :, 1].sum(0), (xy[:, :, 0] >= 0).sum(0).float()], dim=1)
    full = SH.gather_points(local, n, rank, world)
    whole = S.generate_frames(cams, n)
    want = torch.stack([whole[:, :, 0].sum(0), whole[:, :, 1].sum(0), (whole[:, :, 0] >= 0).sum(0).float()], dim=1)
    assert full.shape == (n, 3) and torch.equal(full, want), (n, rank)
    views = (xy[:, :, 0] >= 0).sum(0)
    bad_local = int(torch.nonzero(views < 2)[0]) if bool((views < 2).any()) else -1
    wv = (whole[:, :, 0] >= 0).sum(0)
    want_bad = int(torch.nonzero(wv < 2)[0]) if bool((wv < 2).any()) else -1
    assert SH.first_bad_frame(bad_local, b, "cpu", world) == want_bad
# ---- the frame-sharded classifier's host logic (slice -> enumerate everywhere -> link along the rank chain -> gather)
# with a stand-in engine whose "linking" is a running sum over the frames: any break in the rank order, a lost or
# stale state, or a mis-sliced shard changes the result.  (The CUDA classifier itself: tests/test_gpu_classify.py.)
import numpy as np  # noqa: E402


class ChainProbeEngine:
    cameras = [None, None, None]  # 3 "cameras"
    device = 0

    def classify_begin(self, mode, n_drones, offs, xy, n_frames, flags=0):
        o = np.asarray(offs).reshape(3, n_frames + 1)
        x = np.asarray(xy, np.float64).reshape(-1, 2)
        self.per_frame = np.array([sum(x[o[c, f]:o[c, f + 1]].sum() for c in range(3)) for f in range(n_frames)])
        self.n = (n_drones, n_frames)

    def classify_finish(self, state=None):
        nb = tri_b200.lib().tri_classify_state_bytes()
        run = 0.0 if state is None else float(np.frombuffer(state, np.float64, 1)[0])
        n_drones, n_frames = self.n
        paths = np.zeros((n_drones, n_frames, 3))
        for f in range(n_frames):
            run = run * 0.5 + self.per_frame[f]  # order-dependent recurrence
            paths[:, f, 0] = run
        out = np.zeros(nb, np.uint8)
        out[:8] = np.frombuffer(np.float64(run).tobytes(), np.uint8)
        return dict(paths=paths, assign=np.zeros((n_drones, n_frames, 3), np.int8), phase=np.ones((n_drones, n_frames), np.uint8),
                    stats={"nodes": n_frames, "max_frontier": n_frames}, state=out.tobytes())


rng = np.random.default_rng(5)
nf = 37
cnt = rng.integers(0, 4, size=(3, nf))
offs = np.zeros((3, nf + 1), np.int64)
offs[:, 1:] = np.cumsum(cnt, axis=1)
offs += np.concatenate([[0], np.cumsum(cnt.sum(1))[:-1]])[:, None]
dets = rng.normal(size=(int(cnt.sum()), 2))
r = SH.classify_sharded(ChainProbeEngine(), 0, 2, offs.reshape(-1).astype(np.int32), dets, nf, rank, world)
full = SH.gather_classified(r, nf, rank, world)
solo = ChainProbeEngine()
solo.classify_begin(0, 2, offs.reshape(-1).astype(np.int32), dets, nf)
want = solo.classify_finish(None)
assert np.array_equal(full["paths"], want["paths"]) and full["paths"].shape == (2, nf, 3), rank
assert full["stats"]["nodes"] == nf
dist.barrier()
if rank == 0:
    print("DIST_OK")
dist.destroy_process_group()
