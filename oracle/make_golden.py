"""Generate the golden OUTPUT vectors under tests/golden/ with the Python/cv2 twin (oracle/py_twin.py:
the reference's arithmetic on the real OpenCV 4.13 primitives).  Run HERE (needs cv2); the npz files are
committed and are what pins the C oracle and, through it, the CUDA engine.

  python oracle/make_golden.py [--quick]
"""
import argparse
import itertools
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import oracle_py as O  # noqa: E402  (fixture loaders only)
import py_twin as T  # noqa: E402

G = os.path.join(os.path.dirname(HERE), "tests", "golden")


def twin_cams(name):
    return [T.make_camera(*t) for t in O.parse_cameras_xml(os.path.join(G, name + "_cameras.xml"))]


def nested(offs, xy, n_cam, n_frames):
    o = offs.reshape(n_cam, n_frames + 1)
    return [[[tuple(map(float, v)) for v in xy[o[c, f]:o[c, f + 1]]] for f in range(n_frames)] for c in range(n_cam)]


def classify_golden(cams, mode, n_drones, data, frames, n_cam):
    cl = T.Classifier(cams, mode, n_drones)
    paths, empty = cl.classify(data, frames=frames)
    P = np.array(paths, np.float64).reshape(n_drones, frames, 3)
    assign = np.full((n_drones, frames, n_cam), -1, np.int8)
    phase = np.zeros((n_drones, frames), np.uint8)
    for (f, p, comb, ph) in cl.assign_log:
        assign[p, f] = comb
        phase[p, f] = ph
    return dict(paths=P, assign=assign, phase=phase, ties=np.int64(cl.ties),
                solves=np.int64(cl.counters.get("solves", 0)), nodes=np.int64(cl.counters.get("nodes", 0)),
                leaves=np.int64(cl.counters.get("leaves", 0)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    t0 = time.time()

    cam_out = {}
    for name in ("R02_D1", "S09_D6"):
        cams = twin_cams(name)
        cam_out[name + "_P"] = np.array([c.P for c in cams])
        cam_out[name + "_K"] = np.array([c.K for c in cams])
        cam_out[name + "_E"] = np.array([c.E for c in cams])
        cam_out[name + "_fov"] = np.array([[c.fovx, c.fovy, c.fx, c.fy, c.cx, c.cy] for c in cams])
    np.savez_compressed(G + "/golden_cameras.npz", **cam_out)

    # ---- R02_D1: batch API (K1b, K2f) on getDataForTriangulation() ----
    cams = twin_cams("R02_D1")
    offs, xy, nc, nf = O.load_dets(G + "/R02_D1_dets.npz")
    pts = O.dets_to_points(offs, xy, nc, nf)
    m_xyz, m_err = np.zeros((nf, 3)), np.zeros(nf)
    r_xyz, r_err, r_it = np.zeros((nf, 3)), np.zeros(nf), np.zeros(nf, np.int32)
    for f in range(nf):
        im = [(c, pts[c, f, 0], pts[c, f, 1]) for c in range(nc)]
        X, e = T.matrix_point(cams, im)
        m_xyz[f], m_err[f] = X, e
        st = {}
        X, e = T.ray_point(cams, im, st)
        r_xyz[f], r_err[f], r_it[f] = X, e, st["iters"]
    np.savez_compressed(G + "/golden_R02_D1_batch.npz", matrix_xyz=m_xyz, matrix_err=m_err, ray_xyz=r_xyz,
                        ray_err=r_err, ray_iters=r_it)
    print("R02_D1 batch done", time.time() - t0, "ray iters mean", r_it.mean(), flush=True)

    # ---- R02_D1: every 2- and 3-view subset on a frame sample (K1a, K2e incl. non-converging LM) ----
    frames = list(range(0, nf, 40 if not a.quick else 400))
    rows = []
    for f in frames:
        for k in (2, 3):
            for sub in itertools.combinations(range(nc), k):
                im = [(c, pts[c, f, 0], pts[c, f, 1]) for c in sub]
                Xm, em = T.matrix_point(cams, im)
                st = {}
                Xr, er = T.ray_point(cams, im, st)
                mask = sum(1 << c for c in sub)
                rows.append([f, mask, *Xm, em, *Xr, er, st["iters"]])
    np.savez_compressed(G + "/golden_R02_D1_subsets.npz", rows=np.array(rows, np.float64))
    print("R02_D1 subsets done", time.time() - t0, flush=True)

    # ---- classifier (K3) ----
    data = nested(offs, xy, nc, nf)
    np.savez_compressed(G + "/golden_R02_D1_classify_matrix.npz", **classify_golden(cams, "matrix", 1, data, nf, nc))
    print("R02_D1 classify matrix done", time.time() - t0, flush=True)
    nfr = 12 if a.quick else 40
    np.savez_compressed(G + "/golden_R02_D1_classify_ray.npz", **classify_golden(cams, "ray", 1, data, nfr, nc))
    print("R02_D1 classify ray done", time.time() - t0, flush=True)

    cams = twin_cams("S09_D6")
    offs, xy, nc, nf = O.load_dets(G + "/S09_D6_dets.npz")
    data = nested(offs, xy, nc, nf)
    nfr = 10 if a.quick else 120
    np.savez_compressed(G + "/golden_S09_D6_classify_matrix.npz", **classify_golden(cams, "matrix", 6, data, nfr, nc))
    print("S09_D6 classify matrix done", time.time() - t0, flush=True)


if __name__ == "__main__":
    main()
