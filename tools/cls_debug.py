"""First (frame, path) where tri_classify and the CPU oracle disagree on a fixture, with context.
  python tools/cls_debug.py [--dataset S09_D6] [--drones 6] [--frames N]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import oracle_py as O
import tri_b200 as T
ap = argparse.ArgumentParser()
ap.add_argument("--dataset", default="S09_D6"); ap.add_argument("--drones", type=int, default=6); ap.add_argument("--frames", type=int, default=0)
a = ap.parse_args()
G = os.path.join(ROOT, "tests", "golden")
cams = T.load_cameras_xml("%s/%s_cameras.xml" % (G, a.dataset))
offs, xy, nc, nf = O.load_dets("%s/%s_dets.npz" % (G, a.dataset))
if a.frames: offs, xy, nc, nf = O.slice_frames(offs, xy, nc, nf, 0, a.frames)
oc = [O.make_camera(c.cam_id, c.width, c.height, c.focal, c.position, c.quat) for c in cams]
ref = O.classify(oc, O.MATRIX, a.drones, offs, xy, nc, nf)
eng = T.Engine(cams, 0)
r = eng.classify(T.MATRIX, a.drones, offs, xy, nf)
r = eng.classify(T.MATRIX, a.drones, offs, xy, nf)
print("stats", r["stats"], "oracle", ref["stats"])
bad = np.argwhere((r["assign"] != ref["assign"]).any(axis=2) | (r["phase"] != ref["phase"]))
print("mismatching (path, frame) cells:", len(bad))
if len(bad):
    f = int(bad[:, 1].min())
    print("first frame", f)
    for fr in range(max(0, f - 1), min(nf, f + 2)):
        print(" frame", fr, "detections per camera", [int(offs.reshape(nc, nf + 1)[c, fr + 1] - offs.reshape(nc, nf + 1)[c, fr]) for c in range(nc)])
        for p in range(a.drones):
            print("   path %d  gpu phase %d assign %s pt %s | oracle phase %d assign %s pt %s" % (p, r["phase"][p, fr], r["assign"][p, fr].tolist(), np.round(r["paths"][p, fr], 2).tolist(),
                  ref["phase"][p, fr], ref["assign"][p, fr].tolist(), np.round(ref["paths"][p, fr], 2).tolist()))
