"""Frame-sharded classifier under torchrun (one process per GPU): every rank enumerates its frame range, the shards
are linked in rank order with the tracking state sent rank to rank.
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_classify_sharded.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import oracle_py as O  # noqa: E402  (fixture loader)
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
for name, mode, flags in (("matrix", T.MATRIX, 0), ("ray-closed-form", T.RAY, T.RAY_CLOSED_FORM), ("ray-reference-LM", T.RAY, T.RAY_REFERENCE_LM)):
    whole = eng.classify(mode, 6, offs, xy, nf, flags) if rank == 0 else None  # also the warm-up of rank 0
    SH.classify_sharded(eng, mode, 6, offs, xy, nf, rank, world, flags)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = SH.classify_sharded(eng, mode, 6, offs, xy, nf, rank, world, flags)
    dist.barrier()
    dt = time.perf_counter() - t0
    full = SH.gather_classified(r, nf, rank, world)
    if rank == 0:
        t1 = time.perf_counter()
        eng.classify(mode, 6, offs, xy, nf, flags)
        one = time.perf_counter() - t1
        print(json.dumps({"dataset": "S09_D6", "mode": name, "n_gpus": world, "sharded_s": dt, "one_gpu_s": one,
                          "identical": bool(np.array_equal(full["assign"], whole["assign"]) and np.array_equal(full["paths"], whole["paths"]))}), flush=True)
dist.destroy_process_group()
