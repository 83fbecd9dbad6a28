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
dist.barrier()
if rank == 0:
    print("DIST_OK")
dist.destroy_process_group()
