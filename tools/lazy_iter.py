"""The lazy best-first classifier on a 32-camera rig (config 5): link_us per frame of one sequence, wall time of 128 recordings.
  python tools/lazy_iter.py [--frames N]"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import tri_b200 as T  # noqa: E402
from tri_b200 import synthetic as S  # noqa: E402
from tri_b200 import sharding as SH  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=2000)
a = ap.parse_args()
cams = S.ring_rig(32, rings=((6000.0, 3000.0), (9000.0, 5000.0)))
eng = T.Engine(cams, 0)
q, l = 128, 250
offs, xy, truth = S.generate_multi_drone(cams, q * l, 6)
o1, x1 = SH.slice_csr(offs, xy, 32, q * l, 0, a.frames)
ref = None
for rep in range(3):
    r = eng.classify(T.MATRIX, 6, o1, x1, a.frames)
    if ref is None:
        ref = r
    assert np.array_equal(ref["assign"], r["assign"])
    print("32 cameras x %d frames: link_us %d (%.1f us/frame) nodes %d solves %d" % (a.frames, r["stats"]["link_us"], r["stats"]["link_us"] / a.frames,
                                                                             r["stats"]["nodes"], r["stats"]["solves"]), flush=True)
print("assignment checksum", int(np.asarray(ref["assign"], np.int64).sum()), "paths checksum %.6f" % float(np.nansum(ref["paths"])))
b = np.arange(0, q * l + 1, l, dtype=np.int32)
t0 = time.perf_counter()
r4 = eng.classify_sequences(T.MATRIX, 6, b, offs, xy, q * l)
dt = time.perf_counter() - t0
print("128 recordings x 250 frames: %.3f s, %.0f frames/s, link_us %d nodes %d" % (dt, q * l / dt, r4["stats"]["link_us"], r4["stats"]["nodes"]))
for nfr in (1, 2, 50, 500):
    o2, x2 = SH.slice_csr(offs, xy, 32, q * l, 0, nfr)
    r = eng.classify(T.MATRIX, 6, o2, x2, nfr)
    print("first %d frames: link_us %d nodes %d solves %d phase1 %d phase2 %d" % (nfr, r["stats"]["link_us"], r["stats"]["nodes"], r["stats"]["solves"], r["stats"]["phase1"], r["stats"]["phase2"]))
