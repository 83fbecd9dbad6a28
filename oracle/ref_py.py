"""ctypes binding of oracle/_ref/libref.so -- the REFERENCE'S OWN sources compiled unmodified against oracle/shim
(oracle/Makefile, oracle/ref_harness.cpp).  TEST INFRASTRUCTURE ONLY: tests/, smoke() and bench.py's CPU legs.

`available()` is False where the library has not been built (it needs /root/reference at build time; the built
files travel to the GPU box with the snapshot)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libref.so")
MAIN = os.path.join(HERE, "_ref", "ref_main")
MATRIX, RAY = 0, 1
_lib = None


def available():
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB)
        L.ref_last_error.restype = C.c_char_p
        L.ref_create.restype = C.c_void_p
        L.ref_create_xml.restype = C.c_void_p
        L.ref_create_xml.argtypes = [C.c_char_p, C.c_int]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_n_cameras.argtypes = [C.c_void_p]
        L.ref_camera.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4
        L.ref_triangulate_points.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p]
        L.ref_triangulate_point.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_dist_from_ray.restype = C.c_double
        L.ref_dist_from_ray.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p]
        L.ref_classify.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 6
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Reference:
    """One reference Triangulator (MatrixTriangulator | RayTriangulator) over cameras built by createCamera."""

    def __init__(self, cams=None, mode=MATRIX, xml=None):
        """cams: tuples (id, width, height, focal, pos[3], quat[4]) -- the fields of src/utils.cpp:46-92."""
        self.mode = mode
        if xml is not None:
            self.h = lib().ref_create_xml(xml.encode(), mode)
        else:
            ids = np.array([c[0] for c in cams], np.int32)
            w = np.array([c[1] for c in cams], np.int32)
            h = np.array([c[2] for c in cams], np.int32)
            f = np.array([c[3] for c in cams], np.float64)
            pos = np.array([c[4] for c in cams], np.float64).reshape(-1)
            quat = np.array([c[5] for c in cams], np.float64).reshape(-1)
            self.h = lib().ref_create(len(cams), _p(ids), _p(w), _p(h), _p(f), _p(pos), _p(quat), mode)
        if not self.h:
            raise RuntimeError(lib().ref_last_error().decode())
        self.n_cams = lib().ref_n_cameras(self.h)

    def close(self):
        if self.h:
            lib().ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def camera(self, i):
        P, K, E, s = np.zeros(12), np.zeros(9), np.zeros(12), np.zeros(8)
        if lib().ref_camera(self.h, i, _p(P), _p(K), _p(E), _p(s)):
            raise RuntimeError(lib().ref_last_error().decode())
        return dict(P=P.reshape(3, 4), K=K.reshape(3, 3), E=E.reshape(3, 4), fovx=s[0], fovy=s[1], fx=s[2], fy=s[3],
                    cx=int(s[4]), cy=int(s[5]), width=int(s[6]), height=int(s[7]))

    def triangulate_points(self, xy):
        """xy [n_point_cams][n_frames][2] float64 -> [n_frames][3]; raises RuntimeError with the reference's text."""
        xy = np.ascontiguousarray(xy, np.float64)
        out = np.zeros((xy.shape[1], 3))
        if lib().ref_triangulate_points(self.h, _p(xy), xy.shape[0], xy.shape[1], _p(out)):
            raise RuntimeError(lib().ref_last_error().decode())
        return out

    def triangulate_point(self, cam_idx, xy):
        idx = np.asarray(cam_idx, np.int32)
        pix = np.ascontiguousarray(xy, np.float64).reshape(-1)
        X, err, it = np.zeros(3), C.c_double(0), C.c_int(0)
        if lib().ref_triangulate_point(self.h, len(idx), _p(idx), _p(pix), _p(X), C.byref(err), C.byref(it)):
            raise RuntimeError(lib().ref_last_error().decode())
        return X, err.value, it.value

    def dist_from_ray(self, cam, x, y, p):
        p = np.ascontiguousarray(p, np.float64)
        return lib().ref_dist_from_ray(self.h, cam, x, y, _p(p))

    def classify(self, n_drones, offs, xy, n_frames):
        """DroneClassifier::classifyDrones -> dict(paths, assign, phase, err, stats, margins)."""
        offs = np.ascontiguousarray(offs, np.int32)
        xy = np.ascontiguousarray(xy, np.float64)
        paths = np.zeros((n_drones, n_frames, 3))
        assign = np.zeros((n_drones, n_frames, self.n_cams), np.int8)
        phase = np.zeros((n_drones, n_frames), np.uint8)
        err = np.zeros((n_drones, n_frames))
        stats, margins = np.zeros(4, np.int64), np.zeros(3)
        if lib().ref_classify(self.h, n_drones, _p(offs), _p(xy), n_frames, _p(paths), _p(assign), _p(phase), _p(err),
                              _p(stats), _p(margins)):
            raise RuntimeError(lib().ref_last_error().decode())
        return dict(paths=paths, assign=assign, phase=phase, err=err,
                    stats=dict(solves=int(stats[0]), lm_iters=int(stats[1]), phase1=int(stats[2]), phase2=int(stats[3])),
                    margins=dict(error=float(margins[0]), step=float(margins[1]), gate=float(margins[2])))
