"""Python/cv2 twin of the reference hot path -- TEST INFRASTRUCTURE ONLY.

This file restates the arithmetic of Grzetan/3D-Reconstruction-Triangulation on top of the
in-container OpenCV python wheel (cv2 4.13), so that the third-party pieces the reference leans
on (cv::invert(DECOMP_SVD), cv::solve/invert(DECOMP_EIG), cv::mulTransposed, cv::gemm) are the
real OpenCV ones.  It is used (a) to generate the committed golden vectors under tests/golden/
(see oracle/make_golden.py) and (b) to pin the plain-C oracle (oracle/tri_oracle.c).  Nothing in
the product path may import it.

Citations are file:line under /root/reference/.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

try:  # cv2 exists in this image; the twin is never needed where it does not
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

DBL_EPSILON = 2.220446049250313e-16
FLT_EPSILON = 1.1920928955078125e-07

# src/RayTriangulator.h:9-11
THRESHOLD = 1e-4
MAX_ITERATIONS = 1000
# src/DroneClassifier.h:11-17
MAX_ERROR_MATRIX = 1e5
MAX_ERROR_RAY = 120.0
MAX_STEP = 200.0
MIN_CAMERAS = 2
PATH_TAIL = 3

# src/Camera.h:23-26
RAD_TO_DEG = 57.29577951308232087679
DEG_TO_RAD = 0.01745329251994329576


# --------------------------------------------------------------------------------------
# P0: tdr::Camera (src/Camera.h:78-187, 274-287) built the way createCamera does
# (src/utils.cpp:94-107)
# --------------------------------------------------------------------------------------
@dataclass
class Camera:
    cam_id: int
    width: int
    height: int
    focal: float
    position: np.ndarray  # "tvec" in the reference = world position (utils.cpp:101)
    quat: np.ndarray  # "rquat", (w,i,j,k) raw, never normalised
    cx: int = 0
    cy: int = 0
    fx: float = 0.0
    fy: float = 0.0
    fovx: float = 0.0
    fovy: float = 0.0
    K: np.ndarray = field(default=None)
    E: np.ndarray = field(default=None)
    P: np.ndarray = field(default=None)
    cam_pos: np.ndarray = field(default=None)


def to_rot_matrix(q):
    """Camera.h:274-287 -- uses the RAW quaternion (qTemp is computed and ignored)."""
    a, b, c, d = (float(q[0]), float(q[1]), float(q[2]), float(q[3]))
    return np.array(
        [
            [1 - 2 * (c * c + d * d), 2 * (b * c - a * d), 2 * (b * d + a * c)],
            [2 * (b * c + a * d), 1 - 2 * (b * b + d * d), 2 * (c * d - a * b)],
            [2 * (b * d - a * c), 2 * (c * d + a * b), 1 - 2 * (b * b + c * c)],
        ],
        dtype=np.float64,
    )


def make_camera(cam_id, width, height, focal, position, quat) -> Camera:
    cam = Camera(cam_id, int(width), int(height), float(focal),
                 np.asarray(position, np.float64).reshape(3), np.asarray(quat, np.float64).reshape(4))
    if cam.width == 0 or cam.height == 0:
        raise RuntimeError("Set width and height first.")  # Camera.h:79-80
    # compCxCy, Camera.h:78-83 (C round(): half away from zero)
    cam.cx = int(math.floor(cam.width / 2.0 + 0.5))
    cam.cy = int(math.floor(cam.height / 2.0 + 0.5))
    # compFovx, Camera.h:89-92 (truncated literal 57.2958)
    cam.fx = cam.focal
    if cam.fx == 0:
        raise RuntimeError("Set fx first.")
    cam.fovx = 2 * math.atan(cam.width / (2 * cam.fx)) * 57.2958
    # compFovy, Camera.h:98-105
    cam.fovy = 2.0 * math.atan(math.tan(cam.fovx * 0.5 * DEG_TO_RAD) / (float(cam.width) / float(cam.height))) * RAD_TO_DEG
    # compFxFy, Camera.h:112-117
    cam.fx = (cam.width / 2.0) / math.tan((cam.fovx / 2.0) * DEG_TO_RAD)
    cam.fy = (cam.height / 2.0) / math.tan((cam.fovy / 2.0) * DEG_TO_RAD)
    # compCamPos, Camera.h:167-170: R = toRotMatrix(rquat).inv() (cv::Mat::inv, DECOMP_LU)
    R = cv2.invert(to_rot_matrix(cam.quat), flags=cv2.DECOMP_LU)[1]
    cam.cam_pos = cv2.gemm(R, cam.position.reshape(3, 1), -1.0, None, 0.0).reshape(3)
    # createCamMat, Camera.h:123-128
    cam.K = np.array([[cam.fx, 0, cam.cx], [0, cam.fy, cam.cy], [0, 0, 1]], np.float64)
    # createExtricsicMat, Camera.h:133-153
    cam.E = np.concatenate([R, cam.cam_pos.reshape(3, 1)], axis=1)
    # createPerspectiveMat, Camera.h:159-161
    cam.P = cv2.gemm(cam.K, cam.E, 1.0, None, 0.0)
    return cam


# --------------------------------------------------------------------------------------
# K1a: MatrixTriangulator::triangulatePoint (src/MatrixTriangulator.cpp:3-62)
# --------------------------------------------------------------------------------------
def matrix_point(cams, images):
    """images: list of (cam_index, x, y). Returns ((X,Y,Z), err)."""
    n = len(images)
    A = np.empty((2 * n, 3), np.float64)
    b = np.empty((2 * n, 1), np.float64)
    for i, (ci, x, y) in enumerate(images):
        P = cams[ci].P
        for k in range(3):
            A[2 * i, k] = P[0, k] - x * P[2, k]
            A[2 * i + 1, k] = P[1, k] - y * P[2, k]
        b[2 * i, 0] = x * P[2, 3] - P[0, 3]
        b[2 * i + 1, 0] = y * P[2, 3] - P[1, 3]
    inv = cv2.invert(A, flags=cv2.DECOMP_SVD)[1]  # :53
    X = cv2.gemm(inv, b, 1.0, None, 0.0)  # :54
    e = cv2.gemm(A, X, 1.0, b, -1.0)  # :55
    ss = float(cv2.gemm(e, e, 1.0, None, 0.0, flags=cv2.GEMM_1_T)[0, 0])  # :56-58
    err = math.sqrt(ss / (2 * n))  # :59
    return (float(X[0, 0]), float(X[1, 0]), float(X[2, 0])), err


# --------------------------------------------------------------------------------------
# K2a-c: ray helpers (src/Triangulator.cpp:3-61)
# --------------------------------------------------------------------------------------
def pixel_dir(cam: Camera, x, y):
    """calculateRayDirectionForPixel, Triangulator.cpp:27-44."""
    px = x + 0.5
    py = y + 0.5
    d = 1 / math.tan(cam.fovy * 0.0174533 / 2)
    vx = (float(cam.width) / float(cam.height)) * ((2 * px / float(cam.width)) - 1)
    vy = (2 * py / float(cam.height)) - 1
    vz = d
    # cv::normalize(Vec3d): v * (1/norm(v)); norm = sqrt(sum of squares)
    nrm = math.sqrt(vx * vx + vy * vy + vz * vz)
    s = 1.0 / nrm
    return (vx * s, vy * s, vz * s)


def quat_mul(p, q):
    """cv::Vec4d operator* (Hamilton product, element 0 = w) -- matx.hpp."""
    a1, b1, c1, d1 = p
    a2, b2, c2, d2 = q
    return (a1 * a2 - b1 * b2 - c1 * c2 - d1 * d2,
            a1 * b2 + b1 * a2 + c1 * d2 - d1 * c2,
            a1 * c2 - b1 * d2 + c1 * a2 + d1 * b2,
            a1 * d2 + b1 * c2 - c1 * b2 + d1 * a2)


def rotate_by_quat(v, q):
    """rotatePointByQuaternion, Triangulator.cpp:15-25: q * (0,v) * conj(q)."""
    qq = (float(q[0]), float(q[1]), float(q[2]), float(q[3]))
    t = quat_mul(qq, (0.0, v[0], v[1], v[2]))
    r = quat_mul(t, (qq[0], -qq[1], -qq[2], -qq[3]))
    return (r[1], r[2], r[3])


def make_ray(cam: Camera, x, y):
    """createRayForPoint, Triangulator.cpp:46-55. origin = cam->tvec (the world position)."""
    return (tuple(float(t) for t in cam.position), rotate_by_quat(pixel_dir(cam, x, y), cam.quat))


def dist_to_ray(ray, p):
    """distToRay, Triangulator.cpp:3-9 (squared=false): |dir x (p - origin)|."""
    (ox, oy, oz), (dx, dy, dz) = ray
    wx, wy, wz = p[0] - ox, p[1] - oy, p[2] - oz
    cx = dy * wz - dz * wy
    cy = dz * wx - dx * wz
    cz = dx * wy - dy * wx
    return math.sqrt(cx * cx + cy * cy + cz * cz)


# --------------------------------------------------------------------------------------
# cv::LMSolver (opencv 4.x calib3d/src/levmarq.cpp, LMSolverImpl::run), restated on cv2 primitives
# --------------------------------------------------------------------------------------
def lm_run(compute, x0, max_iters, eps=FLT_EPSILON, stats=None):
    """compute(x, need_J) -> (r (m,1), J (m,n) or None).  Returns (x, iters)."""
    x = np.asarray(x0, np.float64).reshape(-1, 1).copy()
    lx = x.shape[0]
    r, J = compute(x, True)
    S = float(cv2.norm(r, cv2.NORM_L2SQR))
    A = cv2.mulTransposed(J, True)
    v = cv2.gemm(J, r, 1.0, None, 0.0, flags=cv2.GEMM_1_T)
    D = A.diagonal().copy()
    Rlo, Rhi = 0.25, 0.75
    lam, lc = 1.0, 0.75
    it = 0
    while True:
        Ap = A.copy()
        for i in range(lx):
            Ap[i, i] += lam * D[i]
        d = cv2.solve(Ap, v, flags=cv2.DECOMP_EIG)[1]
        xd = x - d
        rd, _ = compute(xd, False)
        Sd = float(cv2.norm(rd, cv2.NORM_L2SQR))
        temp_d = cv2.gemm(A, d, -1.0, v, 2.0)
        dS = float(d.ravel().dot(temp_d.ravel())) if False else float(cv2.gemm(d, temp_d, 1.0, None, 0.0, flags=cv2.GEMM_1_T)[0, 0])
        R = (S - Sd) / (dS if abs(dS) > DBL_EPSILON else 1.0)
        if R > Rhi:
            lam *= 0.5
            if lam < lc:
                lam = 0.0
        elif R < Rlo:
            t = float(cv2.gemm(d, v, 1.0, None, 0.0, flags=cv2.GEMM_1_T)[0, 0])
            nu = (Sd - S) / (t if abs(t) > DBL_EPSILON else 1.0) + 2
            nu = min(max(nu, 2.0), 10.0)
            if lam == 0:
                Ai = cv2.invert(A, flags=cv2.DECOMP_EIG)[1]
                maxval = DBL_EPSILON
                for i in range(lx):
                    maxval = max(maxval, abs(Ai[i, i]))
                lam = lc = 1.0 / maxval
                nu *= 0.5
            lam *= nu
        if Sd < S:
            S = Sd
            x, xd = xd, x
            r, J = compute(x, True)
            A = cv2.mulTransposed(J, True)
            v = cv2.gemm(J, r, 1.0, None, 0.0, flags=cv2.GEMM_1_T)
        it += 1
        proceed = it < max_iters and float(cv2.norm(d, cv2.NORM_INF)) >= eps and float(cv2.norm(r, cv2.NORM_INF)) >= eps
        if not proceed:
            break
    if stats is not None:
        stats["iters"] = it
    return x.reshape(-1), it


# --------------------------------------------------------------------------------------
# K2d/e: RayTriangulator (src/RayTriangulator.cpp:8-47, 83-107)
# --------------------------------------------------------------------------------------
def ray_point(cams, images, stats=None):
    rays = [make_ray(cams[ci], x, y) for (ci, x, y) in images]
    n = len(rays)
    # initGuess: cv::Point3d += then /= size, RayTriangulator.cpp:90-98
    gx = gy = gz = 0.0
    for (o, _) in rays:
        gx += o[0]
        gy += o[1]
        gz += o[2]
    gx /= n
    gy /= n
    gz /= n
    last_err = [0.0]
    eps = THRESHOLD

    def compute(xm, need_J):
        x, y, z = float(xm[0, 0]), float(xm[1, 0]), float(xm[2, 0])
        r = np.empty((n, 1), np.float64)
        s = 0.0
        for i, ray in enumerate(rays):  # sequential order (see race note SURVEY 5)
            e = dist_to_ray(ray, (x, y, z))
            r[i, 0] = e
            s += e
        last_err[0] = s / float(n)
        J = None
        if need_J:
            J = np.empty((n, 3), np.float64)
            for i, ray in enumerate(rays):
                J[i, 0] = (dist_to_ray(ray, (x + eps, y, z)) - dist_to_ray(ray, (x - eps, y, z))) / (2 * eps)
                J[i, 1] = (dist_to_ray(ray, (x, y + eps, z)) - dist_to_ray(ray, (x, y - eps, z))) / (2 * eps)
                J[i, 2] = (dist_to_ray(ray, (x, y, z + eps)) - dist_to_ray(ray, (x, y, z - eps))) / (2 * eps)
        return r, J

    p, it = lm_run(compute, [gx, gy, gz], MAX_ITERATIONS, stats=stats)
    return (float(p[0]), float(p[1]), float(p[2])), last_err[0]


def ray_closed_form(cams, images):
    """Exact minimiser of sum |d x (p-o)|^2 (independent check, not a reference function)."""
    M = np.zeros((3, 3))
    c = np.zeros(3)
    for (ci, x, y) in images:
        o, d = make_ray(cams[ci], x, y)
        d = np.array(d)
        o = np.array(o)
        Mi = d.dot(d) * np.eye(3) - np.outer(d, d)
        M += Mi
        c += Mi @ o
    return np.linalg.solve(M, c)


# --------------------------------------------------------------------------------------
# K1b / K2f: triangulatePoints (MatrixTriangulator.cpp:70-100, RayTriangulator.cpp:51-81)
# --------------------------------------------------------------------------------------
def triangulate_points(cams, points, mode):
    """points[cam][frame] = (x,y) with (-1,-1) sentinel."""
    for i in range(len(points) - 1):
        if len(points[i]) != len(points[i + 1]):
            raise RuntimeError("Every camera should have the same number of points")
    out = []
    ncam = min(len(points), len(cams)) if mode == "matrix" else len(points)
    for f in range(len(points[0])):
        images = []
        for c in range(ncam):
            x, y = points[c][f]
            if x == -1 or y == -1:
                continue
            images.append((c, x, y))
        if len(images) < 2:
            raise RuntimeError("Too few rays are found" if mode == "matrix" else "Too few detections are found")
        fn = matrix_point if mode == "matrix" else ray_point
        out.append(fn(cams, images)[0])
    return out


# --------------------------------------------------------------------------------------
# K4b: DetectionsContainer::readFiles (src/DetectionsContainer.cpp:19-76)
# --------------------------------------------------------------------------------------
def _stoi(tok: str) -> int:
    s = tok.lstrip(" \t\n\v\f\r")
    i = 0
    if i < len(s) and s[i] in "+-":
        i += 1
    j = i
    while j < len(s) and s[j].isdigit():
        j += 1
    if j == i:
        raise ValueError("stoi")
    return int(s[:j])


def read_csv_files(files, offset=0, record_size=7):
    data = []
    for f in files:
        cam = []
        frame = -1
        n_line = 0
        with open(f, "r", newline="") as fh:
            text = fh.read()
        lines = text.split("\n")
        if lines and lines[-1] == "":
            lines.pop()
        for line in lines:
            if offset > n_line:
                n_line += 1
                continue
            n_line += 1
            sep = [_stoi(t) for t in line.split(",") if t != ""]
            for _ in range(sep[0] - frame - 1):
                cam.append([])
            frame = sep[0]
            if (len(sep) - 1) % record_size != 0:
                raise RuntimeError("Invalid CSV file!")
            dets = []
            for j in range(len(sep) // record_size):
                dets.append((float(sep[j * record_size + 5]), float(sep[j * record_size + 6])))
            cam.append(dets)
        data.append(cam)
    if len(data) < 2:
        raise RuntimeError("There must be at least 2 cameras")
    for i in range(1, len(data)):
        if len(data[i - 1]) != len(data[i]):
            raise RuntimeError("Number of frames on all cameras must be the same")
    return data


# --------------------------------------------------------------------------------------
# K3: DroneClassifier (src/DroneClassifier.cpp)
# --------------------------------------------------------------------------------------
class _Iterator:
    """DroneClassifier::Iterator, DroneClassifier.cpp:43-89."""

    def __init__(self, sizes):
        self.sizes = list(sizes)
        self.c = [-1] * len(sizes)
        self.skip = False

    def increment(self):
        if self.skip:
            self.skip = False
            return True
        for i in range(len(self.c)):
            if self.c[i] == -1:
                self.c[i] += 1
                return True
        for i in range(len(self.c) - 1, -1, -1):
            if self.c[i] < self.sizes[i] - 1:
                self.c[i] += 1
                return True
            self.c[i] = -1
        return False

    def cut(self):
        for i in range(len(self.c) - 1, -1, -1):
            if self.c[i] != -1 and self.c[i] < self.sizes[i] - 1:
                self.c[i] += 1
                self.skip = True
                return True
            self.c[i] = -1
        return False


@dataclass
class Combination:
    comb: tuple
    point: tuple
    error: float
    order: int = 0  # DFS enumeration order (tie audit only)

    def zeros(self):
        return sum(1 for v in self.comb if v == 0)


def _unique(c, combos):
    """Combination::isCombinationUnique, DroneClassifier.cpp:32-41."""
    for o in combos:
        for i in range(len(c.comb)):
            if c.comb[i] == o.comb[i] and c.comb[i] != 0:
                return False
    return True


class Classifier:
    def __init__(self, cams, mode, n_drones, counters=None):
        self.cams = cams
        self.mode = mode
        self.n_drones = n_drones
        self.error_ = MAX_ERROR_MATRIX if mode == "matrix" else MAX_ERROR_RAY
        self.counters = counters if counters is not None else {}
        self.ties = 0
        self.assign_log = []  # (frame, path, comb, phase)

    def _tri(self, images):
        self.counters["solves"] = self.counters.get("solves", 0) + 1
        return matrix_point(self.cams, images) if self.mode == "matrix" else ray_point(self.cams, images)

    def fill_queue(self, dets, remap=None):
        """fillCombinationQueue, DroneClassifier.cpp:156-198. dets[cam] = list of (x,y).
        remap[cam][k] = original combination index of local index k (getOriginalCombination)."""
        sizes = [len(d) + 1 for d in dets]
        it = _Iterator(sizes)
        out = []
        while it.increment():
            comb = list(it.c)
            self.counters["nodes"] = self.counters.get("nodes", 0) + 1
            if sum(1 for v in comb if v > 0) < 2:
                continue
            images = [(i, dets[i][v - 1][0], dets[i][v - 1][1]) for i, v in enumerate(comb) if v > 0]
            pt, err = self._tri(images)
            if err > self.error_:
                if not it.cut():
                    break
            elif -1 not in comb and sum(1 for v in comb if v > 0) >= MIN_CAMERAS:
                orig = tuple(comb) if remap is None else tuple(remap[i][v] for i, v in enumerate(comb))
                out.append(Combination(orig, pt, err, len(out)))
        self.counters["leaves"] = self.counters.get("leaves", 0) + len(out)
        # std::priority_queue order: fewer zeros first, then smaller error (operator<, :12-20).
        srt = sorted(out, key=lambda c: (c.zeros(), c.error, c.order))
        for a, b in zip(srt, srt[1:]):
            if a.zeros() == b.zeros() and a.error == b.error:
                self.ties += 1
        return srt

    def with_last_pos(self, pos, frame_dets, used):
        """triangulateWithLastPos, DroneClassifier.cpp:219-250."""
        gated = []
        remap = []
        for cam, dets in enumerate(frame_dets):
            g = []
            m = {0: 0}
            for det, (x, y) in enumerate(dets):
                if dist_to_ray(make_ray(self.cams[cam], x, y), pos) < MAX_STEP:
                    g.append((x, y))
                    m[len(g)] = det + 1
            gated.append(g)
            remap.append(m)
        for c in self.fill_queue(gated, remap):
            dx, dy, dz = c.point[0] - pos[0], c.point[1] - pos[1], c.point[2] - pos[2]
            if _unique(c, used) and c.error < self.error_ and math.sqrt(dx * dx + dy * dy + dz * dz) < MAX_STEP:
                return c
        return None

    def pick_best(self, frame_dets, used):
        """pickBestCombinations, DroneClassifier.cpp:200-217."""
        final = []
        for c in self.fill_queue(frame_dets):
            if _unique(c, final) and _unique(c, used) and c.error < self.error_:
                final.append(c)
        return final

    def classify_paths(self, final, paths, processed, empty_frames, frame):
        """classifyPaths, DroneClassifier.cpp:262-332."""
        cpv = []
        for i, fc in enumerate(final):
            best_path, best_dist = 0, -1.0
            for j in range(self.n_drones):
                if j in processed:
                    continue
                n = min(len(paths[j]), PATH_TAIL)
                if n == 0:
                    continue
                dist = 0.0
                for t in range(len(paths[j]) - n, len(paths[j])):
                    q = paths[j][t]
                    dx, dy, dz = q[0] - fc.point[0], q[1] - fc.point[1], q[2] - fc.point[2]
                    dist += math.sqrt(dx * dx + dy * dy + dz * dz)
                dist /= float(n)
                if dist < best_dist or best_dist == -1:
                    best_dist, best_path = dist, j
            cpv.append((i, best_path, best_dist))
        # std::sort(greater<>) with operator> = (error < elem.error): ascending error.
        # libstdc++ uses insertion sort for <= 16 elements (stable); we define stable always.
        cpv.sort(key=lambda e: e[2])
        for (ci, path, _) in cpv:
            if path in processed:
                empty = -1
                for i in range(len(paths)):
                    if len(paths[i]) == 0:
                        empty = i
                        break
                if empty != -1:
                    paths[empty].append(final[ci].point)
                    processed.append(empty)
                    self.assign_log.append((frame, empty, final[ci].comb, 2))
            else:
                paths[path].append(final[ci].point)
                processed.append(path)
                self.assign_log.append((frame, path, final[ci].comb, 2))
        for i in range(self.n_drones):
            if i not in processed:
                empty_frames[i].append(frame)

    def classify(self, data, frames=None):
        """classifyDrones, DroneClassifier.cpp:96-154. data[cam][frame] = list of (x,y)."""
        n_frames = len(data[0]) if frames is None else frames
        paths = [[] for _ in range(self.n_drones)]
        empty_frames = [[] for _ in range(self.n_drones)]
        for frame in range(n_frames):
            frame_dets = [data[c][frame] for c in range(len(data))]
            processed = []
            used = []
            for n_path in range(len(paths)):
                cur = paths[n_path]
                if len(cur) == 0:
                    continue
                if cur[-1] != (0.0, 0.0, 0.0):
                    best = self.with_last_pos(cur[-1], frame_dets, used)
                    if best is not None:
                        processed.append(n_path)
                        used.append(best)
                        paths[n_path].append(best.point)
                        self.assign_log.append((frame, n_path, best.comb, 1))
            if len(processed) == len(paths):
                continue
            final = self.pick_best(frame_dets, used)
            self.classify_paths(final, paths, processed, empty_frames, frame)
        for i in range(len(empty_frames)):
            for j in empty_frames[i]:
                paths[i].insert(j, (0.0, 0.0, 0.0))
        return paths, empty_frames
