"""world_size-2 NCCL worker for tests/test_gpu_classify.py: frame-sharded classification, one process per GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py as O  # noqa: E402  (fixture loader only)
import tri_b200 as T  # noqa: E402
from tri_b200 import sharding as SH  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
G = os.path.join(ROOT, "tests", "golden")
cams = T.load_cameras_xml(G + "/S09_D6_cameras.xml")
offs, xy, nc, nf = O.load_dets(G + "/S09_D6_dets.npz")
eng = T.Engine(cams, local)
r = SH.classify_sharded(eng, T.MATRIX, 6, offs, xy, nf, rank, world)
full = SH.gather_classified(r, nf, rank, world)
if rank == 0:
    whole = eng.classify(T.MATRIX, 6, offs, xy, nf)
    assert np.array_equal(full["assign"], whole["assign"]) and np.array_equal(full["phase"], whole["phase"])
    assert np.array_equal(full["paths"], whole["paths"])
    print("DIST_CLASSIFY_OK")
dist.barrier()
dist.destroy_process_group()
