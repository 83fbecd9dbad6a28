"""Multi-GPU host logic on the CPU: frame-range sharding and the final gather (gloo, world_size 2)."""
import os
import subprocess
import sys

import tri_b200  # noqa: F401
from tri_b200 import sharding as SH

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_partition_the_frames():
    for n in (0, 1, 2, 3, 17, 1000, 1001, 100_000_000):
        for world in (1, 2, 3, 4, 8):
            r = [SH.shard_range(n, g, world) for g in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[g][1] == r[g + 1][0] for g in range(world - 1))
            assert all(b <= e for b, e in r)
            assert all(b % 2 == 0 for b, e in r)  # shard starts stay vector-aligned
            if n >= 2 * world:
                sizes = [e - b for b, e in r]
                assert max(sizes) - min(sizes) <= 2


def test_gather_world_size_2_gloo():
    port = 29500 + os.getpid() % 2000
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "_dist_worker.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DIST_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_slice_csr_matches_the_oracle_slicer():
    """Frame ranges of the classifier's CSR detections (what each rank of the sharded classifier enumerates):
    same offsets / pixels as the oracle's own slicer, on a real multi-drone fixture with ragged detection counts."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as O
    offs, xy, nc, nf = O.load_dets(os.path.join(ROOT, "tests", "golden", "S09_D6_dets.npz"))
    for f0, f1 in ((0, nf), (0, 1), (17, 18), (100, 1600), (nf - 3, nf), (5, 5)):
        o, x = SH.slice_csr(offs, xy, nc, nf, f0, f1)
        wo, wx, wnc, wnf = O.slice_frames(offs, xy, nc, nf, f0, f1)
        assert wnf == f1 - f0 and np.array_equal(o, np.asarray(wo).reshape(-1))
        assert np.array_equal(np.asarray(x).reshape(-1, 2), np.asarray(wx).reshape(-1, 2))
    # the ranges classify_chain / classify_sharded use partition the sequence
    for world in (1, 2, 3, 8):
        cuts = [(nf * g // world, nf * (g + 1) // world) for g in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == nf and all(cuts[g][1] == cuts[g + 1][0] for g in range(world - 1))
