"""Synthetic rigs and detections for the throughput configs of BASELINE.json.

Everything is keyed by the GLOBAL frame index through a counter-based hash (splitmix64), so any
sharding of the frame range over GPUs sees the same data.  Runs in torch on whatever device the
caller names (the bench generates on the GPU; the parity tests copy the same tensors to the host for
the oracle).
"""
from __future__ import annotations

import math

import numpy as np

from . import Camera

SEED = 20240607


def look_at_quat(pos, target=(0.0, 0.0, 0.0)):
    """(w,x,y,z) of the camera-to-world rotation whose +z axis points from pos to target, x to the
    right and y down -- the convention of the reference's ray model (Triangulator.cpp:15-55) and of
    Camera::createExtricsicMat (Camera.h:133-153: extrinsic R = toRotMatrix(q)^-1)."""
    p, t = np.asarray(pos, float), np.asarray(target, float)
    z = (t - p) / np.linalg.norm(t - p)
    up = np.array([0.0, 0.0, 1.0])
    x = np.cross(z, up)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    R = np.stack([x, y, z], axis=1)  # columns = camera axes in world coordinates
    t_ = R[0, 0] + R[1, 1] + R[2, 2]
    if t_ > 0:
        s = math.sqrt(t_ + 1.0) * 2
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
    return tuple(float(v) for v in q)


def ring_rig(n_cams=8, rings=((6000.0, 3000.0),), width=1920, height=1080, focal=1081.0810546875):
    """Cameras evenly spaced on ring(s) (radius, z) in mm, looking at the origin; built through the
    same Camera arithmetic as the real rigs.  BASELINE config 4: 8 cameras, one ring; config 5: 32
    cameras on two rings ((6000,3000),(9000,5000))."""
    cams = []
    per = n_cams // len(rings)
    for r_i, (radius, z) in enumerate(rings):
        n = per if r_i < len(rings) - 1 else n_cams - per * (len(rings) - 1)
        for k in range(n):
            a = 2 * math.pi * (k + 0.5 * r_i) / n
            pos = (radius * math.cos(a), radius * math.sin(a), z)
            cams.append(Camera(len(cams) + 1, width, height, focal, pos, look_at_quat(pos, (0.0, 0.0, 1000.0))))
    return cams


def _hash01(torch, idx, stream):
    """splitmix64(idx, stream) -> float64 uniform in (0,1); idx int64 tensor."""
    M34, M37, M33 = (1 << 34) - 1, (1 << 37) - 1, (1 << 33) - 1
    z = idx + (SEED + 0x632BE59BD9B4E019 * (stream + 1)) % (1 << 63)
    z = z * -7046029254386353131  # 0x9E3779B97F4A7C15 as int64
    z = (z ^ ((z >> 30) & M34)) * -4658895280553007687  # 0xBF58476D1CE4E5B9
    z = (z ^ ((z >> 27) & M37)) * -7723592293110705685  # 0x94D049BB133111EB
    z = z ^ ((z >> 31) & M33)
    return (((z >> 11) & ((1 << 53) - 1)).to(torch.float64) + 0.5) * (1.0 / (1 << 53))


def generate_frames(cams, n_frames, frame0=0, device="cpu", p_missing=0.2, noise_px=1.0, out=None, chunk=1 << 22,
                    volume=((-2000.0, 2000.0), (-2000.0, 2000.0), (200.0, 2500.0)), want_truth=False, dtype=None):
    """-> xy [n_cams, n_frames, 2] float32 (integer-valued pixels, (-1,-1) = missing) on `device`
    (+ truth [n_frames, 3] float64 if asked).  Pixels: project with the reference's own P, add
    N(0, noise_px), round to the nearest integer (real detections are integers,
    DetectionsContainer.cpp:36), then drop each (camera, frame) with probability p_missing."""
    import torch
    dtype = dtype or torch.float32
    nc = len(cams)
    if out is None:
        out = torch.empty((nc, n_frames, 2), dtype=dtype, device=device)
    truth = torch.empty((n_frames, 3), dtype=torch.float64, device=device) if want_truth else None
    P = torch.tensor(np.stack([c.P for c in cams]), dtype=torch.float64, device=device)  # [nc,3,4]
    for a in range(0, n_frames, chunk):
        n = min(chunk, n_frames - a)
        idx = torch.arange(frame0 + a, frame0 + a + n, dtype=torch.int64, device=device)
        X = torch.stack([lo + (hi - lo) * _hash01(torch, idx, k) for k, (lo, hi) in enumerate(volume)], dim=1)
        if truth is not None:
            truth[a:a + n] = X
        Xh = torch.cat([X, torch.ones((n, 1), dtype=torch.float64, device=device)], dim=1)  # [n,4]
        for c in range(nc):
            h = Xh @ P[c].T  # [n,3]
            u1, u2 = _hash01(torch, idx, 16 + 4 * c), _hash01(torch, idx, 17 + 4 * c)
            rad = torch.sqrt(-2.0 * torch.log(u1)) * noise_px
            px = torch.round(h[:, 0] / h[:, 2] + rad * torch.cos(2 * math.pi * u2))
            py = torch.round(h[:, 1] / h[:, 2] + rad * torch.sin(2 * math.pi * u2))
            miss = _hash01(torch, idx, 18 + 4 * c) < p_missing
            if dtype == torch.uint16:
                px = torch.where(miss, torch.full_like(px, 65535.0), px.clamp(0, 65534))
                py = torch.where(miss, torch.full_like(py, 65535.0), py.clamp(0, 65534))
                out[c, a:a + n, 0] = px.to(torch.int32).to(torch.uint16)
                out[c, a:a + n, 1] = py.to(torch.int32).to(torch.uint16)
            else:
                px = torch.where(miss, torch.full_like(px, -1.0), px)
                py = torch.where(miss, torch.full_like(py, -1.0), py)
                out[c, a:a + n, 0] = px.to(dtype)
                out[c, a:a + n, 1] = py.to(dtype)
    return (out, truth) if want_truth else out
