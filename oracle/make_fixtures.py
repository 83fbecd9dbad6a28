"""Build the input fixtures under tests/golden/ from the reference's datasets -- run HERE only
(needs /root/reference); the outputs are committed so that nothing reads /root/reference at test time.

  python oracle/make_fixtures.py

Produces
  tests/golden/R02_D1_cameras.xml   cameras.xml in the format src/utils.cpp:46-92 parses, from
                                    dataset/R02_D1/R02_D1.xcp (ORIENTATION x y z w -> w -x -y -z)
  tests/golden/S09_D6_cameras.xml   same, from dataset/S09_D6/stationary_camera_data.csv
                                    (the recipe of cameraDataConverter.py without Blender mathutils)
  tests/golden/R02_D1_dets.npz      detections as parsed by DetectionsContainer::readFiles
  tests/golden/S09_D6_dets.npz      (src/DetectionsContainer.cpp:19-76): counts[cam,frame] uint8,
                                    xy int16 (n_det,2) in [cam][frame][det] order
  tests/golden/csv_sample/*.csv     first rows of two S09_D6 files + the rows around a frame with no
                                    detections, verbatim, for the CSV-reader tests
"""
import glob
import math
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import py_twin as T  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def xcp_cameras(path):
    txt = open(path).read()
    cams = []
    for m in re.finditer(r"<Camera\b([^>]*)>(.*?)</Camera>", txt, re.S):
        head, body = m.group(1), m.group(2)
        cf = re.search(r"<ControlFrame\b([^>]*)/>", body)
        if not cf:
            continue
        at = dict(re.findall(r'(\w+)="([^"]*)"', cf.group(1)))
        dev = re.search(r'DEVICEID="(\d+)"', head).group(1)
        x, y, z, w = at["ORIENTATION"].split()
        neg = lambda s: s[1:] if s.startswith("-") else "-" + s  # exact textual negation
        cams.append(dict(id=dev, focal=at["FOCAL_LENGTH"], pp=at["PRINCIPAL_POINT"], pos=at["POSITION"],
                         ori=" ".join([w, neg(x), neg(y), neg(z)])))
    return cams


def rodrigues(rv):
    th = math.sqrt(rv[0] ** 2 + rv[1] ** 2 + rv[2] ** 2)
    k = np.array(rv) / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + math.sin(th) * Kx + (1 - math.cos(th)) * (Kx @ Kx)


def mat_to_quat(R):
    """Rotation matrix -> (w,x,y,z), Shepperd's method."""
    t = R[0, 0] + R[1, 1] + R[2, 2]
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        q = (0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s)
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = math.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = ((R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s)
    elif R[1, 1] > R[2, 2]:
        s = math.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = ((R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s)
    else:
        s = math.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = ((R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s)
    return q


def stationary_cameras(path):
    cams = []
    for i, line in enumerate(open(path).read().splitlines()[1:]):
        f = line.split(";")
        if len(f) < 10 or not f[0]:
            continue
        t = np.array([float(f[1]), float(f[2]), float(f[3])])
        R = rodrigues([float(f[4]), float(f[5]), float(f[6])])
        pos = -(R.T @ t) * 1000.0  # mm: the thresholds of DroneClassifier.h:11-15 are in mm
        q = mat_to_quat(R.T)
        fov, w, h = float(f[7]), int(f[8]), int(f[9])
        focal = (w / 2) / math.tan(math.radians(fov) / 2)
        cams.append(dict(id=str(i + 1), focal=repr(focal), pp="%d %d" % (w // 2, h // 2),
                         pos=" ".join(repr(float(v)) for v in pos), ori=" ".join(repr(float(v)) for v in q)))
    return cams


def write_xml(cams, path):
    with open(path, "w") as f:
        f.write('<?xml version="1.0" encoding="UTF-8"?>\n<Cameras>\n')
        for c in cams:
            f.write('  <Camera DEVICEID="%s">\n    <ControlFrames>\n' % c["id"])
            f.write('      <ControlFrame FOCAL_LENGTH="%s" FRAME="0" ORIENTATION="%s" POSITION="%s" PRINCIPAL_POINT="%s"/>\n'
                    % (c["focal"], c["ori"], c["pos"], c["pp"]))
            f.write("    </ControlFrames>\n  </Camera>\n")
        f.write("</Cameras>\n")


def pack_dets(files, path):
    data = T.read_csv_files(files)
    n_cam, n_frames = len(data), len(data[0])
    counts = np.zeros((n_cam, n_frames), np.uint8)
    xy = []
    for c in range(n_cam):
        for f in range(n_frames):
            counts[c, f] = len(data[c][f])
            xy.extend(data[c][f])
    xy = np.array(xy, np.float64)
    assert np.all(xy == np.round(xy)) and np.abs(xy).max() < 32000
    np.savez_compressed(path, counts=counts, xy=xy.astype(np.int16),
                        files=np.array([os.path.basename(f) for f in files]))
    print(path, counts.shape, xy.shape, "max dets", counts.max())


def main():
    os.makedirs(OUT, exist_ok=True)
    write_xml(xcp_cameras(REF + "/dataset/R02_D1/R02_D1.xcp"), OUT + "/R02_D1_cameras.xml")
    write_xml(stationary_cameras(REF + "/dataset/S09_D6/stationary_camera_data.csv"), OUT + "/S09_D6_cameras.xml")
    pack_dets(sorted(glob.glob(REF + "/dataset/R02_D1/dl_data/*.csv")), OUT + "/R02_D1_dets.npz")
    pack_dets(sorted(glob.glob(REF + "/dataset/S09_D6/dl_data/*.csv")), OUT + "/S09_D6_dets.npz")
    os.makedirs(OUT + "/csv_sample", exist_ok=True)
    for name in ("cam5i_s.csv", "cam6i_s.csv"):
        raw = open(REF + "/dataset/S09_D6/dl_data/" + name, "rb").read().split(b"\n")
        open(OUT + "/csv_sample/" + name, "wb").write(b"\n".join(raw[:40]) + b"\n")


if __name__ == "__main__":
    main()
