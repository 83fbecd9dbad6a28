"""More fixtures + golden vectors (two-drone datasets) generated HERE from /root/reference with the cv2 twin:
  R04_D2   real rig (4 cameras, .xcp), 2 drones, first 1000 frames of 14128
  S01_D2_A simulated rig (8 cameras), 2 drones, 500 frames
Run: python oracle/make_golden_extra.py"""
import glob
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_fixtures as MF  # noqa: E402
import make_golden as MG  # noqa: E402
import oracle_py as O  # noqa: E402
import py_twin as T  # noqa: E402

REF, OUT = MF.REF, MF.OUT


def pack(files, path, n_frames):
    data = T.read_csv_files(files)
    data = [cam[:n_frames] for cam in data]
    n_cam, nf = len(data), len(data[0])
    counts = np.zeros((n_cam, nf), np.uint8)
    xy = []
    for c in range(n_cam):
        for f in range(nf):
            counts[c, f] = len(data[c][f])
            xy.extend(data[c][f])
    xy = np.array(xy, np.float64)
    np.savez_compressed(path, counts=counts, xy=xy.astype(np.int16), files=np.array([os.path.basename(f) for f in files]))
    print(path, counts.shape, xy.shape, "max dets", counts.max())


def main():
    t0 = time.time()
    MF.write_xml(MF.xcp_cameras(REF + "/dataset/R04_D2/R04_D2.xcp"), OUT + "/R04_D2_cameras.xml")
    MF.write_xml(MF.stationary_cameras(REF + "/dataset/S01_D2_A/stationary_camera_data.csv"), OUT + "/S01_D2_A_cameras.xml")
    pack(sorted(glob.glob(REF + "/dataset/R04_D2/dl_data/*.csv")), OUT + "/R04_D2_dets.npz", 1000)
    pack(sorted(glob.glob(REF + "/dataset/S01_D2_A/dl_data/*.csv")), OUT + "/S01_D2_A_dets.npz", 500)
    for name, mode, frames in (("R04_D2", "matrix", 300), ("S01_D2_A", "matrix", 100), ("R04_D2", "ray", 16)):
        cams = MG.twin_cams(name)
        offs, xy, nc, nf = O.load_dets(OUT + "/%s_dets.npz" % name)
        data = MG.nested(offs, xy, nc, nf)
        np.savez_compressed(OUT + "/golden_%s_classify_%s.npz" % (name, mode), **MG.classify_golden(cams, mode, 2, data, frames, nc))
        print(name, mode, frames, "done", time.time() - t0, flush=True)


if __name__ == "__main__":
    main()
