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
