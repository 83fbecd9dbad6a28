"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY (see tri_oracle.h).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  Also holds the fixture loaders (cameras.xml / dets.npz under tests/golden/).
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")

MATRIX, RAY = 0, 1
OK, ERR_DIM, ERR_TOO_FEW, ERR_CAMERA = 0, 1, 2, 3


class OrcCamera(C.Structure):
    _fields_ = [("id", C.c_int), ("width", C.c_int), ("height", C.c_int), ("cx", C.c_int), ("cy", C.c_int),
                ("focal", C.c_double), ("fx", C.c_double), ("fy", C.c_double), ("fovx", C.c_double),
                ("fovy", C.c_double), ("pos", C.c_double * 3), ("quat", C.c_double * 4),
                ("cam_pos", C.c_double * 3), ("K", C.c_double * 9), ("E", C.c_double * 12), ("P", C.c_double * 12)]


class OrcStats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("nodes", "solves", "leaves", "lm_iters", "ties", "phase1", "phase2")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def build():
    subprocess.check_call(["make", "-s", "-C", HERE])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.orc_matrix_point.restype = C.c_double
        _lib.orc_ray_point.restype = C.c_double
        _lib.orc_dist_from_ray.restype = C.c_double
        _lib.orc_dist_to_ray.restype = C.c_double
    return _lib


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def make_camera(cam_id, width, height, focal, pos, quat):
    c = OrcCamera()
    st = lib().orc_camera_make(C.byref(c), int(cam_id), int(width), int(height), C.c_double(focal),
                               (C.c_double * 3)(*pos), (C.c_double * 4)(*quat))
    if st != OK:
        raise RuntimeError("camera parameters rejected (Camera.h:79-80/90/124)")
    return c


def camera_array(cams):
    arr = (OrcCamera * len(cams))()
    for i, c in enumerate(cams):
        arr[i] = c
    return arr


def parse_cameras_xml(path):
    """The fields src/utils.cpp:46-92 reads, as plain tuples (id, w, h, focal, pos, quat)."""
    out = []
    txt = open(path).read()
    for m in re.finditer(r"<Camera\b([^>]*)>(.*?)</Camera>", txt, re.S):
        cf = re.search(r"<ControlFrame\b([^>]*)/>", m.group(2))
        if not cf:
            continue
        at = dict(re.findall(r'(\w+)="([^"]*)"', cf.group(1)))
        dev = int(re.search(r'DEVICEID="(-?\d+)"', m.group(1)).group(1))
        w, h = (2 * v for v in _istream_two_ints(at["PRINCIPAL_POINT"]))  # utils.cpp:64-71
        out.append((dev, w, h, float(at["FOCAL_LENGTH"]), [float(v) for v in at["POSITION"].split()],
                    [float(v) for v in at["ORIENTATION"].split()]))
    return out


def _istream_two_ints(text):
    """`std::stringstream(text) >> a >> b` on ints: a failed extraction yields 0 and stops."""
    vals, pos = [0, 0], 0
    for k in range(2):
        m = re.match(r"\s*([+-]?\d+)", text[pos:])
        if not m:
            break
        vals[k] = int(m.group(1))
        pos += m.end()
    return vals


def load_cameras(path):
    return [make_camera(*t) for t in parse_cameras_xml(path)]


def load_dets(path):
    """-> (offsets int32 [n_cam*(n_frames+1)], xy float64 [n,2], n_cam, n_frames)."""
    z = np.load(path)
    counts = z["counts"].astype(np.int64)
    n_cam, n_frames = counts.shape
    offs = np.zeros((n_cam, n_frames + 1), np.int64)
    base = 0
    for c in range(n_cam):
        offs[c, 0] = base
        offs[c, 1:] = base + np.cumsum(counts[c])
        base = offs[c, -1]
    return offs.astype(np.int32).reshape(-1).copy(), z["xy"].astype(np.float64).copy(), n_cam, n_frames


def dets_to_points(offs, xy, n_cam, n_frames):
    """DetectionsContainer::getDataForTriangulation (DetectionsContainer.cpp:145-171): [cam][frame][2]
    with the (-1,-1) sentinel; raises like the reference if a (cam,frame) has >1 detection."""
    o = offs.reshape(n_cam, n_frames + 1)
    cnt = o[:, 1:] - o[:, :-1]
    if cnt.max() > 1:
        raise RuntimeError("Function 'getDataForTriangulation' can be used only for one drone. "
                           "Each frame can have max one detection")
    pts = np.full((n_cam, n_frames, 2), -1.0)
    c, f = np.nonzero(cnt)
    pts[c, f] = xy[o[c, f]]
    return pts


def matrix_point(cams, cam_idx, xy):
    arr = camera_array(cams)
    idx = np.asarray(cam_idx, np.int32)
    pix = np.asarray(xy, np.float64).reshape(-1)
    X = np.zeros(3)
    err = lib().orc_matrix_point(arr, len(idx), _p(idx, C.c_int), _p(pix), _p(X))
    return X, err


def ray_point(cams, cam_idx, xy):
    arr = camera_array(cams)
    idx = np.asarray(cam_idx, np.int32)
    pix = np.asarray(xy, np.float64).reshape(-1)
    X = np.zeros(3)
    it = C.c_int(0)
    err = lib().orc_ray_point(arr, len(idx), _p(idx, C.c_int), _p(pix), _p(X), C.byref(it))
    return X, err, it.value


def ray_closed_form(cams, cam_idx, xy):
    arr = camera_array(cams)
    idx = np.asarray(cam_idx, np.int32)
    pix = np.asarray(xy, np.float64).reshape(-1)
    X = np.zeros(3)
    lib().orc_ray_closed_form(arr, len(idx), _p(idx, C.c_int), _p(pix), _p(X))
    return X


def make_ray(cam, x, y):
    o, d = np.zeros(3), np.zeros(3)
    lib().orc_make_ray(C.byref(cam), C.c_double(x), C.c_double(y), _p(o), _p(d))
    return o, d


def triangulate_points(cams, xy, mode, allow_too_few=False, nthreads=1, want_iters=False):
    """xy: [n_point_cams][n_frames][2] float64 or float32. -> dict(xyz, err, mask, iters, status)."""
    arr = camera_array(cams)
    xy = np.ascontiguousarray(xy)
    npc, nf = xy.shape[0], xy.shape[1]
    out = np.zeros((nf, 3))
    err = np.zeros(nf)
    mask = np.zeros(nf, np.uint32)
    iters = np.zeros(nf, np.int32) if want_iters else None
    if xy.dtype == np.float32:
        fn, pt = lib().orc_triangulate_points_f32, _p(xy, C.c_float)
    else:
        xy = xy.astype(np.float64, copy=False)
        fn, pt = lib().orc_triangulate_points, _p(xy)
    st = fn(arr, len(cams), npc, mode, pt, C.c_int64(nf), int(allow_too_few), _p(out), _p(err),
            _p(mask, C.c_uint32), _p(iters, C.c_int32) if want_iters else None, int(nthreads))
    return dict(xyz=out, err=err, mask=mask, iters=iters, status=st)


def classify(cams, mode, n_drones, offs, xy, n_cam, n_frames):
    arr = camera_array(cams)
    paths = np.zeros((n_drones, n_frames, 3))
    assign = np.zeros((n_drones, n_frames, n_cam), np.int8)
    phase = np.zeros((n_drones, n_frames), np.uint8)
    st = OrcStats()
    offs = np.ascontiguousarray(offs, np.int32)
    xy = np.ascontiguousarray(xy, np.float64)
    rc = lib().orc_classify(arr, n_cam, mode, n_drones, _p(offs, C.c_int32), _p(xy), n_frames, _p(paths),
                            _p(assign, C.c_int8), _p(phase, C.c_uint8), C.byref(st))
    if rc != OK:
        raise RuntimeError("orc_classify status %d" % rc)
    m = np.zeros(5)
    lib().orc_last_margins(_p(m))
    return dict(paths=paths, assign=assign, phase=phase, stats=st.as_dict(),
                margins=dict(error=m[0], step=m[1], gate=m[2], order=m[3], tail=m[4]))


def enumerate_frame(cams, mode, offs, xy, n_cam, n_frames, frame, max_leaves=1 << 16):
    arr = camera_array(cams)
    comb = np.zeros((max_leaves, n_cam), np.int8)
    pts = np.zeros((max_leaves, 3))
    err = np.zeros(max_leaves)
    st = OrcStats()
    offs = np.ascontiguousarray(offs, np.int32)
    xy = np.ascontiguousarray(xy, np.float64)
    n = lib().orc_enumerate_frame(arr, n_cam, mode, _p(offs, C.c_int32), _p(xy), n_frames, frame, max_leaves,
                                  _p(comb, C.c_int8), _p(pts), _p(err), C.byref(st))
    n = min(n, max_leaves)
    return dict(comb=comb[:n], xyz=pts[:n], err=err[:n], stats=st.as_dict())


def slice_frames(offs, xy, n_cam, n_frames, f0, f1):
    """Detections of frames [f0,f1) as a fresh CSR (same [cam][frame][det] order)."""
    o = offs.reshape(n_cam, n_frames + 1)
    new_offs = np.zeros((n_cam, f1 - f0 + 1), np.int32)
    parts = []
    base = 0
    for c in range(n_cam):
        a, b = o[c, f0], o[c, f1]
        new_offs[c] = o[c, f0:f1 + 1] - a + base
        parts.append(xy[a:b])
        base += b - a
    return new_offs.reshape(-1).copy(), np.concatenate(parts, axis=0).copy(), n_cam, f1 - f0
