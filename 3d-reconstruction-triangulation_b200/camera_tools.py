"""Camera-file tooling: produce the `cameras.xml` that loadCamerasXML
(src/utils.cpp:46-92 of the reference) parses, from the two calibration sources the datasets ship.

* Vicon `.xcp` (real datasets R02_D1 ...): the `<Camera>` nodes that carry a `<ControlFrame>`; the XCP's
  ORIENTATION is `x y z w`, the loader reads `w i j k`, and the stored rotation is the inverse of what
  Camera.h expects, so ORIENTATION becomes `w -x -y -z`.  Replaces convert_xcp.py.
* `stationary_camera_data.csv` (simulated datasets S09_D6 ...): OpenCV rvec/tvec in metres; position =
  -R^T t in mm (the classifier thresholds of DroneClassifier.h:11-15 are mm), orientation = quaternion of
  R^T, focal from the horizontal field of view.  Replaces cameraDataConverter.py (no Blender mathutils).

  python -m tri_b200.camera_tools xcp  R02_D1.xcp  cameras.xml
  python -m tri_b200.camera_tools csv  stationary_camera_data.csv  cameras.xml
"""
from __future__ import annotations

import math
import re
import sys


def _neg(s):
    return s[1:] if s.startswith("-") else "-" + s  # exact textual negation, no float round trip


def cameras_from_xcp(path):
    txt = open(path).read()
    cams = []
    for m in re.finditer(r"<Camera\b([^>]*)>(.*?)</Camera>", txt, re.S):
        cf = re.search(r"<ControlFrame\b([^>]*)/>", m.group(2))
        if not cf:
            continue
        at = dict(re.findall(r'(\w+)="([^"]*)"', cf.group(1)))
        dev = re.search(r'DEVICEID="(\d+)"', m.group(1)).group(1)
        x, y, z, w = at["ORIENTATION"].split()
        cams.append(dict(id=dev, focal=at["FOCAL_LENGTH"], pp=at["PRINCIPAL_POINT"], pos=at["POSITION"],
                         ori=" ".join([w, _neg(x), _neg(y), _neg(z)])))
    return cams


def _rodrigues(rv):
    th = math.sqrt(rv[0] ** 2 + rv[1] ** 2 + rv[2] ** 2)
    k = [v / th for v in rv]
    K = [[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]]
    KK = [[sum(K[i][t] * K[t][j] for t in range(3)) for j in range(3)] for i in range(3)]
    s, c = math.sin(th), 1 - math.cos(th)
    return [[(1.0 if i == j else 0.0) + s * K[i][j] + c * KK[i][j] for j in range(3)] for i in range(3)]


def _quat_of(R):
    """Rotation matrix -> (w,x,y,z), Shepperd's method."""
    t = R[0][0] + R[1][1] + R[2][2]
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        return (0.25 * s, (R[2][1] - R[1][2]) / s, (R[0][2] - R[2][0]) / s, (R[1][0] - R[0][1]) / s)
    if R[0][0] > R[1][1] and R[0][0] > R[2][2]:
        s = math.sqrt(1.0 + R[0][0] - R[1][1] - R[2][2]) * 2
        return ((R[2][1] - R[1][2]) / s, 0.25 * s, (R[0][1] + R[1][0]) / s, (R[0][2] + R[2][0]) / s)
    if R[1][1] > R[2][2]:
        s = math.sqrt(1.0 + R[1][1] - R[0][0] - R[2][2]) * 2
        return ((R[0][2] - R[2][0]) / s, (R[0][1] + R[1][0]) / s, 0.25 * s, (R[1][2] + R[2][1]) / s)
    s = math.sqrt(1.0 + R[2][2] - R[0][0] - R[1][1]) * 2
    return ((R[1][0] - R[0][1]) / s, (R[0][2] + R[2][0]) / s, (R[1][2] + R[2][1]) / s, 0.25 * s)


def cameras_from_stationary_csv(path):
    cams = []
    for i, line in enumerate(open(path).read().splitlines()[1:]):
        f = line.split(";")
        if len(f) < 10 or not f[0]:
            continue
        t = [float(f[1]), float(f[2]), float(f[3])]
        R = _rodrigues([float(f[4]), float(f[5]), float(f[6])])
        Rt = [[R[j][i_] for j in range(3)] for i_ in range(3)]
        pos = [-(Rt[a][0] * t[0] + Rt[a][1] * t[1] + Rt[a][2] * t[2]) * 1000.0 for a in range(3)]
        q = _quat_of(Rt)
        fov, w, h = float(f[7]), int(f[8]), int(f[9])
        focal = (w / 2) / math.tan(math.radians(fov) / 2)
        cams.append(dict(id=str(i + 1), focal=repr(focal), pp="%d %d" % (w // 2, h // 2),
                         pos=" ".join(repr(float(v)) for v in pos), ori=" ".join(repr(float(v)) for v in q)))
    return cams


def write_cameras_xml(cams, path):
    with open(path, "w") as f:
        f.write('<?xml version="1.0" encoding="UTF-8"?>\n<Cameras>\n')
        for c in cams:
            f.write('  <Camera DEVICEID="%s">\n    <ControlFrames>\n' % c["id"])
            f.write('      <ControlFrame FOCAL_LENGTH="%s" FRAME="0" ORIENTATION="%s" POSITION="%s" PRINCIPAL_POINT="%s"/>\n'
                    % (c["focal"], c["ori"], c["pos"], c["pp"]))
            f.write("    </ControlFrames>\n  </Camera>\n")
        f.write("</Cameras>\n")


def main(argv):
    if len(argv) != 4 or argv[1] not in ("xcp", "csv"):
        print(__doc__)
        return 2
    cams = cameras_from_xcp(argv[2]) if argv[1] == "xcp" else cameras_from_stationary_csv(argv[2])
    write_cameras_xml(cams, argv[3])
    print("%d cameras -> %s" % (len(cams), argv[3]))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
