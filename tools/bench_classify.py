"""DroneClassifier::classifyDrones on the GPU vs the CPU oracle (datasets of BASELINE configs 1-3).
  python tools/bench_classify.py [--frames N] [--skip-cpu]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402
import oracle_py as O  # noqa: E402  (CPU baseline / checker)
import tri_b200 as T  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--skip-cpu", action="store_true")
ap.add_argument("--ray-frames", type=int, default=200)
a = ap.parse_args()
G = os.path.join(ROOT, "tests", "golden")
res = []
for name, mode, flags, n_drones, frames in (("R02_D1", T.MATRIX, 0, 1, None), ("S09_D6", T.MATRIX, 0, 6, None),
                                            ("R02_D1", T.RAY, T.RAY_REFERENCE_LM, 1, a.ray_frames), ("S09_D6", T.RAY, T.RAY_CLOSED_FORM, 6, None),
                                            ("S09_D6", T.RAY, T.RAY_REFERENCE_LM, 6, None)):
    cams = T.load_cameras_xml(G + "/%s_cameras.xml" % name)
    offs, xy, nc, nf = O.load_dets(G + "/%s_dets.npz" % name)
    if frames:
        offs, xy, nc, nf = O.slice_frames(offs, xy, nc, nf, 0, frames)
    eng = T.Engine(cams, 0)
    eng.classify(mode, n_drones, offs, xy, nf, flags)  # warm-up (allocations, module load)
    t0 = time.perf_counter()
    r = eng.classify(mode, n_drones, offs, xy, nf, flags)
    gpu_s = time.perf_counter() - t0
    row = {"dataset": name, "mode": "matrix" if mode == T.MATRIX else ("ray-reference-LM" if flags & T.RAY_REFERENCE_LM else "ray-closed-form"),
           "n_drones": n_drones, "frames": nf, "gpu_s": gpu_s, "gpu_frames_per_s": nf / gpu_s, "stats": r["stats"]}
    if not a.skip_cpu and not (mode == T.RAY and (flags & T.RAY_CLOSED_FORM or name == "S09_D6")):  # the CPU oracle needs hours for S09_D6 ray
        oc = [O.make_camera(c.cam_id, c.width, c.height, c.focal, c.position, c.quat) for c in cams]
        t0 = time.perf_counter()
        ref = O.classify(oc, O.MATRIX if mode == T.MATRIX else O.RAY, n_drones, offs, xy, nc, nf)
        cpu_s = time.perf_counter() - t0
        row.update(cpu_s=cpu_s, speedup=cpu_s / gpu_s, cpu_solves=ref["stats"]["solves"],
                   assign_equal=bool(np.array_equal(ref["assign"], r["assign"])), phase_equal=bool(np.array_equal(ref["phase"], r["phase"])))
    res.append(row)
    print(json.dumps(row), flush=True)
