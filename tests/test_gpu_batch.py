"""GPU parity of the batched triangulatePoints kernels against the CPU oracle and the golden vectors.
All calls go through the C ABI (tri_b200.Engine -> libtri_b200.so).  Tolerances (BASELINE.json):
validity masks bit-exact; points within 1e-6 relative (FP64) or 1e-4 of scene extent (FP32)."""
import os

import numpy as np
import pytest

import oracle_py as O
import tri_b200 as T
from tri_b200 import synthetic as S

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EXTENT = 1e4  # mm: both rigs span ~10 m


def ocams(cams):
    return [O.make_camera(c.cam_id, c.width, c.height, c.focal, c.position, c.quat) for c in cams]


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available()
    return torch


@pytest.fixture(scope="module")
def r02(torch):
    cams = T.load_cameras_xml(G + "/R02_D1_cameras.xml")
    offs, xy, nc, nf = O.load_dets(G + "/R02_D1_dets.npz")
    pts = O.dets_to_points(offs, xy, nc, nf)
    return cams, T.Engine(cams, 0), pts, np.load(G + "/golden_R02_D1_batch.npz")


@pytest.fixture(scope="module")
def syn(torch):
    cams = S.ring_rig(8)
    xy = S.generate_frames(cams, 200001, device="cuda:0")
    return cams, T.Engine(cams, 0), xy, xy.cpu().numpy()


def rel_err(got, ref):
    return (np.abs(got - ref) / np.maximum(np.abs(ref).max(axis=1, keepdims=True), 1.0)).max()


def test_matrix_fp64_golden_r02(r02, torch):
    cams, eng, pts, g = r02
    out = eng.triangulate_points(T.MATRIX, pts.astype(np.float32), want=("xyz_f64", "err", "mask"))
    assert rel_err(out["xyz_f64"], g["matrix_xyz"]) < 1e-9
    np.testing.assert_allclose(out["err"], g["matrix_err"], rtol=1e-9)
    assert np.all(out["mask"] == 0xF)
    # double2 pixels (cv::Point2d as is) and the device-resident entry give the same
    o2 = eng.triangulate_points(T.MATRIX, pts, want=("xyz_f64",))
    assert np.array_equal(o2["xyz_f64"], out["xyz_f64"])
    d = eng.triangulate_points_device(T.MATRIX, torch.tensor(pts, dtype=torch.float32, device="cuda:0"), want=("xyz_f64", "xyz_f32"))
    assert eng.device_status()[0] == T.OK
    assert np.array_equal(d["xyz_f64"].cpu().numpy(), out["xyz_f64"])
    assert np.abs(d["xyz_f32"].cpu().numpy() - out["xyz_f64"]).max() < 2e-4


def test_matrix_fp32_within_scene_tolerance(r02, syn):
    cams, eng, pts, g = r02
    out = eng.triangulate_points(T.MATRIX, pts.astype(np.float32), T.F32, want=("xyz_f32", "err"))
    assert np.abs(out["xyz_f32"] - g["matrix_xyz"]).max() < 1e-4 * EXTENT
    np.testing.assert_allclose(out["err"], g["matrix_err"], rtol=1e-3)
    cams, eng, xy, host = syn
    ref = O.triangulate_points(ocams(cams), host, O.MATRIX, allow_too_few=True, nthreads=8)
    o = eng.triangulate_points_device(T.MATRIX, xy, T.F32 | T.ALLOW_TOO_FEW, want=("xyz_f32", "mask"))
    eng.device_status()
    assert np.array_equal(o["mask"].cpu().numpy().view(np.uint32), ref["mask"])
    d = np.abs(o["xyz_f32"].cpu().numpy() - ref["xyz"]).max()
    assert d < 1e-4 * EXTENT, d
    assert d < 0.05  # in practice ~1e-6 of the extent


def test_matrix_fp64_synthetic_masks_and_points(syn):
    cams, eng, xy, host = syn
    ref = O.triangulate_points(ocams(cams), host, O.MATRIX, allow_too_few=True, nthreads=8)
    o = eng.triangulate_points_device(T.MATRIX, xy, T.ALLOW_TOO_FEW, want=("xyz_f64", "xyz_f32", "mask", "err"))
    st, bad = eng.device_status()
    few = np.array([bin(int(m)).count("1") < 2 for m in ref["mask"]])
    assert (st == T.ERR_TOO_FEW) == bool(few.any())
    if few.any():
        assert bad == int(np.nonzero(few)[0][0])
    assert np.array_equal(o["mask"].cpu().numpy().view(np.uint32), ref["mask"])  # bit-exact
    got = o["xyz_f64"].cpu().numpy()
    assert rel_err(got, ref["xyz"]) < 1e-9
    assert np.all(got[few] == 0)
    np.testing.assert_allclose(o["err"].cpu().numpy(), ref["err"], rtol=1e-9, atol=1e-9)
    assert np.abs(o["xyz_f32"].cpu().numpy() - got).max() < 1e-3


def test_host_path_equals_device_path_and_throws_like_the_reference(syn):
    cams, eng, xy, host = syn
    o = eng.triangulate_points_device(T.MATRIX, xy, T.ALLOW_TOO_FEW, want=("xyz_f64", "mask"))
    eng.device_status()
    h = eng.triangulate_points(T.MATRIX, host, T.ALLOW_TOO_FEW, want=("xyz_f64", "xyz_f32", "mask"))
    assert np.array_equal(h["xyz_f64"], o["xyz_f64"].cpu().numpy())
    assert np.array_equal(h["mask"], o["mask"].cpu().numpy().view(np.uint32))
    bad = host[:, :4096].copy()
    bad[1:, 77] = -1
    bad[:, 300] = -1
    with pytest.raises(T.TriError) as ei:  # MatrixTriangulator.cpp:92-94
        eng.triangulate_points(T.MATRIX, bad)
    assert ei.value.status == T.ERR_TOO_FEW and str(ei.value) == "Too few rays are found"
    with pytest.raises(T.TriError) as ei:  # RayTriangulator.cpp:72-74
        eng.triangulate_points(T.RAY, bad)
    assert str(ei.value) == "Too few detections are found"
    ok = eng.triangulate_points(T.MATRIX, bad, T.ALLOW_TOO_FEW, want=("xyz_f64", "mask"))
    assert ok["first_bad_frame"] == 77 and ok["mask"][77] == 1 and ok["mask"][300] == 0
    assert np.all(ok["xyz_f64"][[77, 300]] == 0)


@pytest.mark.parametrize("n", [0, 1, 2, 3, 511, 512, 513, 1027])
def test_ragged_sizes_and_unaligned_rows(syn, torch, n):
    cams, eng, xy, host = syn
    ref = O.triangulate_points(ocams(cams), host[:, :n].copy(), O.MATRIX, allow_too_few=True)
    for off in (0, 1):  # off = 1: camera rows start 8 bytes off a 16-byte boundary -> scalar path
        sub = xy[:, off:off + n]
        r = O.triangulate_points(ocams(cams), host[:, off:off + n].copy(), O.MATRIX, allow_too_few=True)
        o = eng.triangulate_points_device(T.MATRIX, sub, T.ALLOW_TOO_FEW, want=("xyz_f64", "xyz_f32", "mask"), n_frames=n)
        eng.device_status()
        assert np.array_equal(o["mask"].cpu().numpy().view(np.uint32), r["mask"])
        if n:
            assert rel_err(o["xyz_f64"].cpu().numpy(), r["xyz"]) < 1e-9
            assert np.abs(o["xyz_f32"].cpu().numpy() - r["xyz"]).max() < 1e-3
    h = eng.triangulate_points(T.MATRIX, host[:, :n].copy(), T.ALLOW_TOO_FEW, want=("xyz_f64",))
    if n:
        assert rel_err(h["xyz_f64"], ref["xyz"]) < 1e-9


def test_camera_count_rules(r02):
    cams, eng, pts, g = r02
    # matrix: min(points.size(), cameras.size()) rows are used (MatrixTriangulator.cpp:84)
    extra = np.concatenate([pts, pts[:1]], axis=0)
    o = eng.triangulate_points(T.MATRIX, extra, want=("xyz_f64",))
    assert rel_err(o["xyz_f64"], g["matrix_xyz"]) < 1e-9
    with pytest.raises(T.TriError) as ei:  # ray would index past the cameras (RayTriangulator.cpp:65-69)
        eng.triangulate_points(T.RAY, extra)
    assert ei.value.status == T.ERR_DIM
    three = eng.triangulate_points(T.MATRIX, pts[:3], want=("xyz_f64", "mask"))
    ref = O.triangulate_points(ocams(cams), pts[:3].copy(), O.MATRIX)
    assert rel_err(three["xyz_f64"], ref["xyz"]) < 1e-9 and np.all(three["mask"] == 7)
    with pytest.raises(T.TriError):
        eng.triangulate_points(T.MATRIX, pts[:1])  # one camera: every frame has too few views


def test_pixel_formats_agree(syn, torch):
    cams, eng, xy, host = syn
    n = 50000
    base = eng.triangulate_points_device(T.MATRIX, xy[:, :n].contiguous(), T.ALLOW_TOO_FEW, want=("xyz_f64", "mask"))
    d64 = eng.triangulate_points_device(T.MATRIX, xy[:, :n].double().contiguous(), T.ALLOW_TOO_FEW, want=("xyz_f64", "mask"))
    u16 = xy[:, :n].clone()
    u16[u16 < 0] = 65535
    u16 = u16.to(torch.int32).to(torch.uint16).contiguous()
    d16 = eng.triangulate_points_device(T.MATRIX, u16, T.ALLOW_TOO_FEW, want=("xyz_f64", "mask"))
    f16 = eng.triangulate_points_device(T.MATRIX, u16, T.ALLOW_TOO_FEW | T.F32, want=("xyz_f32", "mask"))
    eng.device_status()
    for o in (d64, d16):
        assert torch.equal(o["xyz_f64"], base["xyz_f64"]) and torch.equal(o["mask"], base["mask"])
    assert torch.equal(f16["mask"], base["mask"])
    assert (f16["xyz_f32"].double() - base["xyz_f64"]).abs().max() < 0.05


def test_ray_analytic_and_closed_form(r02, syn):
    cams, eng, pts, g = r02
    oc = ocams(cams)
    conv = g["ray_iters"] < 1000
    for flags in (T.RAY_ANALYTIC_LM, 0):
        o = eng.triangulate_points(T.RAY, pts.astype(np.float32), flags, want=("xyz_f64", "err", "iters", "mask"))
        # the reference's LM lands within ~1e-5 mm of the true minimiser where it converges
        assert np.abs(o["xyz_f64"][conv] - g["ray_xyz"][conv]).max() < 1e-3
        assert rel_err(o["xyz_f64"][conv], g["ray_xyz"][conv]) < 1e-6
        np.testing.assert_allclose(o["err"][conv], g["ray_err"][conv], rtol=1e-6)
        for f in range(0, len(pts[0]), 53):
            Xc = O.ray_closed_form(oc, range(4), pts[:, f])
            assert np.abs(o["xyz_f64"][f] - Xc).max() < 1e-7
        assert np.all(o["iters"] <= (4 if flags else 1))
    cams, eng, xy, host = syn
    n = 20000
    a = eng.triangulate_points_device(T.RAY, xy[:, :n].contiguous(), T.ALLOW_TOO_FEW | T.RAY_ANALYTIC_LM, want=("xyz_f64", "mask", "iters"))
    c = eng.triangulate_points_device(T.RAY, xy[:, :n].contiguous(), T.ALLOW_TOO_FEW, want=("xyz_f64", "mask"))
    f = eng.triangulate_points_device(T.RAY, xy[:, :n].contiguous(), T.ALLOW_TOO_FEW | T.F32, want=("xyz_f32",))
    eng.device_status()
    ref = O.triangulate_points(ocams(cams), host[:, :n].copy(), O.RAY, allow_too_few=True, nthreads=8, want_iters=True)
    assert np.array_equal(a["mask"].cpu().numpy().view(np.uint32), ref["mask"])
    ok = (ref["iters"] < 1000) & (ref["iters"] > 0)
    assert ok.mean() > 0.8
    assert np.abs(a["xyz_f64"].cpu().numpy()[ok] - ref["xyz"][ok]).max() < 1e-3
    assert (a["xyz_f64"] - c["xyz_f64"]).abs().max() < 1e-6
    assert (f["xyz_f32"].double() - c["xyz_f64"]).abs().max() < 1e-4 * EXTENT
    assert int(a["iters"].max()) <= 5


def test_ray_reference_lm_follows_the_oracle_trajectory(r02, syn):
    """TRI_RAY_REFERENCE_LM: same iteration counts and bit-identical points/errors as the oracle's
    restatement of cv::LMSolver (which is pinned to real OpenCV through the golden vectors)."""
    cams, eng, pts, g = r02
    ref = O.triangulate_points(ocams(cams), pts, O.RAY, want_iters=True, nthreads=8)
    o = eng.triangulate_points(T.RAY, pts.astype(np.float32), T.RAY_REFERENCE_LM, want=("xyz_f64", "err", "iters"))
    assert np.array_equal(o["iters"], ref["iters"])
    assert np.array_equal(o["xyz_f64"], ref["xyz"])
    assert np.array_equal(o["err"], ref["err"])
    same = o["iters"] == g["ray_iters"]
    assert same.mean() > 0.99
    assert np.abs(o["xyz_f64"][same] - g["ray_xyz"][same]).max() < 1e-6
    cams, eng, xy, host = syn
    n = 3000
    ref = O.triangulate_points(ocams(cams), host[:, :n].copy(), O.RAY, allow_too_few=True, nthreads=8, want_iters=True)
    o = eng.triangulate_points_device(T.RAY, xy[:, :n].contiguous(), T.ALLOW_TOO_FEW | T.RAY_REFERENCE_LM, want=("xyz_f64", "iters", "mask"))
    eng.device_status()
    assert np.array_equal(o["iters"].cpu().numpy(), ref["iters"])
    assert np.array_equal(o["xyz_f64"].cpu().numpy(), ref["xyz"])


def test_subsets_and_dist_from_ray(r02):
    cams, eng, pts, g = r02
    oc = ocams(cams)
    rows = np.load(G + "/golden_R02_D1_subsets.npz")["rows"]
    items = []
    for row in rows:
        f, mask = int(row[0]), int(row[1])
        sub = [c for c in range(4) if mask >> c & 1]
        items.append((sub, pts[sub, f]))
    xyz, err, _ = eng.triangulate_subsets(T.MATRIX, items)
    assert rel_err(xyz, rows[:, 2:5]) < 1e-9
    np.testing.assert_allclose(err, rows[:, 5], rtol=1e-9)
    xyz, err, it = eng.triangulate_subsets(T.RAY, items, T.RAY_REFERENCE_LM)
    for k, (sub, p) in enumerate(items):
        Xo, eo, io = O.ray_point(oc, sub, p)
        assert it[k] == io and np.array_equal(xyz[k], Xo) and err[k] == eo
    xyz_c, err_c, _ = eng.triangulate_subsets(T.RAY, items)
    for k, (sub, p) in enumerate(items[::7]):
        assert np.abs(xyz_c[7 * k] - O.ray_closed_form(oc, sub, p)).max() < 1e-6
    with pytest.raises(T.TriError):
        eng.triangulate_subsets(T.MATRIX, [([0], [[1.0, 2.0]])])
    # getDistFromRay: bit-identical to the oracle (it feeds the < MAX_STEP gate of the classifier)
    rng = np.random.default_rng(3)
    cam = rng.integers(0, 4, 500)
    f = rng.integers(0, pts.shape[1], 500)
    xy = pts[cam, f]
    P = g["matrix_xyz"][f] + rng.normal(0, 150, (500, 3))
    d = eng.dist_from_ray(cam, xy, P)
    want = np.array([O.lib().orc_dist_from_ray(O.C.byref(oc[c]), O.C.c_double(x), O.C.c_double(y), O._p(np.ascontiguousarray(p)))
                     for c, (x, y), p in zip(cam, xy, P)])
    assert np.array_equal(d, want)


def test_full_size_properties(torch):
    """BASELINE config 4 at full size (8 cameras x 100 M frames, 20 % missing): mask statistics,
    agreement with the oracle on a random sample of frames, shard invariance."""
    cams = S.ring_rig(8)
    eng = T.Engine(cams, 0)
    n = 100_000_000
    xy = S.generate_frames(cams, n, device="cuda:0")
    o = eng.triangulate_points_device(T.MATRIX, xy, T.ALLOW_TOO_FEW, want=("xyz_f32", "mask"))
    st, bad = eng.device_status()
    mask = o["mask"]
    pop = torch.zeros_like(mask)
    for c in range(8):
        pop += (mask >> c) & 1
    assert abs(float(pop.double().mean()) - 6.4) < 1e-3
    few = pop < 2
    n_few = int(few.sum())
    assert abs(n_few / n - 8.448e-5) < 1e-5  # P(<2 of 8 | p=0.8)
    assert st == T.ERR_TOO_FEW and bad == int(torch.nonzero(few)[0])
    assert bool((o["xyz_f32"][few] == 0).all())
    idx = torch.randint(0, n, (40000,), device="cuda:0", generator=torch.Generator("cuda:0").manual_seed(1))
    sample = xy[:, idx].cpu().numpy()
    ref = O.triangulate_points(ocams(cams), sample, O.MATRIX, allow_too_few=True, nthreads=8)
    got = o["xyz_f32"][idx].cpu().numpy()
    assert np.array_equal(mask[idx].cpu().numpy().view(np.uint32), ref["mask"])
    assert np.abs(got - ref["xyz"]).max() < 1e-3  # float3 output of the FP64 solve
    # FP32 arithmetic on the same frames
    o32 = eng.triangulate_points_device(T.MATRIX, xy, T.ALLOW_TOO_FEW | T.F32, want=("xyz_f32",))
    eng.device_status()
    assert float((o32["xyz_f32"] - o["xyz_f32"]).abs().max()) < 0.05
    # a shard [a,b) of the frame range gives the same points as the whole
    a, b = 33_333_334, 33_433_334
    sh = eng.triangulate_points_device(T.MATRIX, xy[:, a:b], T.ALLOW_TOO_FEW, want=("xyz_f32",), n_frames=b - a)
    eng.device_status()
    assert torch.equal(sh["xyz_f32"], o["xyz_f32"][a:b])


def test_multi_engine_host_batch(syn, torch):
    """tri_triangulate_points_multi: the batch sharded over every GPU of the box (two engines on the same
    GPU when there is only one) equals the single-engine result bit for bit."""
    cams, eng, xy, host = syn
    n_dev = torch.cuda.device_count()
    engines = [T.Engine(cams, g % n_dev) for g in range(max(2, min(n_dev, 4)))]
    for n in (200001, 1001, 3):
        one = eng.triangulate_points(T.MATRIX, host[:, :n].copy(), T.ALLOW_TOO_FEW, want=("xyz_f64", "mask"))
        many = T.triangulate_points_multi(engines, T.MATRIX, host[:, :n].copy(), T.ALLOW_TOO_FEW, want=("xyz_f64", "mask"))
        assert np.array_equal(one["xyz_f64"], many["xyz_f64"]) and np.array_equal(one["mask"], many["mask"])
        assert one["first_bad_frame"] == many["first_bad_frame"]
    bad = host[:, :5000].copy()
    bad[1:, 4321] = -1
    with pytest.raises(T.TriError) as ei:
        T.triangulate_points_multi(engines, T.MATRIX, bad)
    assert ei.value.status == T.ERR_TOO_FEW


@pytest.mark.parametrize("n_cams", [2, 3, 5, 12, 32])
def test_other_camera_counts(torch, n_cams):
    """Rigs from 2 to TRI_MAX_CAMS cameras (BASELINE config 5 uses 32): unrolled (<= 8) and generic kernels."""
    rings = ((6000.0, 3000.0),) if n_cams <= 8 else ((6000.0, 3000.0), (9000.0, 5000.0))
    cams = S.ring_rig(n_cams, rings=rings)
    eng = T.Engine(cams, 0)
    n = 40001
    xy = S.generate_frames(cams, n, device="cuda:0", p_missing=0.3)
    host = xy.cpu().numpy()
    oc = ocams(cams)
    ref = O.triangulate_points(oc, host, O.MATRIX, allow_too_few=True, nthreads=8)
    o = eng.triangulate_points_device(T.MATRIX, xy, T.ALLOW_TOO_FEW, want=("xyz_f64", "mask", "err"))
    eng.device_status()
    assert np.array_equal(o["mask"].cpu().numpy().view(np.uint32), ref["mask"])
    assert rel_err(o["xyz_f64"].cpu().numpy(), ref["xyz"]) < 1e-8
    ok = ref["err"] > 0
    np.testing.assert_allclose(o["err"].cpu().numpy()[ok], ref["err"][ok], rtol=1e-8, atol=1e-6)
    for fl in (0, T.F32):
        q = eng.triangulate_points_device(T.MATRIX, xy, T.ALLOW_TOO_FEW | fl, want=("xyz_f32", "mask"))
        eng.device_status()
        assert np.array_equal(q["mask"].cpu().numpy().view(np.uint32), ref["mask"])
        assert np.abs(q["xyz_f32"].cpu().numpy() - ref["xyz"]).max() < (1e-3 if fl == 0 else 0.2)
    r = eng.triangulate_points_device(T.RAY, xy, T.ALLOW_TOO_FEW | T.RAY_ANALYTIC_LM, want=("xyz_f64", "mask", "iters"))
    c = eng.triangulate_points_device(T.RAY, xy, T.ALLOW_TOO_FEW, want=("xyz_f64",))
    eng.device_status()
    assert np.array_equal(r["mask"].cpu().numpy().view(np.uint32), ref["mask"])
    assert float((r["xyz_f64"] - c["xyz_f64"]).abs().max()) < 1e-6
    sel = np.nonzero(np.array([bin(int(m)).count("1") for m in ref["mask"]]) >= 3)[0][:300]
    for f in sel[::10]:
        sub = [k for k in range(n_cams) if ref["mask"][f] >> k & 1]
        Xc = O.ray_closed_form(oc, sub, host[sub, f].astype(np.float64))
        assert np.abs(c["xyz_f64"][f].cpu().numpy() - Xc).max() < 1e-6
    # the float3-only outputs take the streaming kernels (<= 8 cameras: table tiles; more: the camera-chunked
    # pipeline) -- same points as the generic kernel's FP64 output, same masks
    for fl, tol in ((T.RAY_ANALYTIC_LM, 1e-3), (0, 1e-3), (T.F32, 0.2)):
        q = eng.triangulate_points_device(T.RAY, xy, T.ALLOW_TOO_FEW | fl, want=("xyz_f32", "mask"))
        eng.device_status()
        assert np.array_equal(q["mask"].cpu().numpy().view(np.uint32), ref["mask"])
        assert float((q["xyz_f32"].double() - c["xyz_f64"]).abs().max()) < tol


def test_host_path_multi_chunk_pipeline(torch):
    """More frames than one staging chunk (1 Mi frames): the three-slot H2D / kernel / D2H pipeline returns
    exactly what one device-resident launch returns, for every output, in both solvers' default precisions."""
    cams = S.ring_rig(8)
    eng = T.Engine(cams, 0)
    n = 2_621_447  # 2.5 chunks, odd
    xy = S.generate_frames(cams, n, device="cuda:0")
    host = xy.cpu().numpy()
    for mode, flags in ((T.MATRIX, 0), (T.MATRIX, T.F32), (T.RAY, 0), (T.RAY, T.RAY_CLOSED_FORM | T.F32)):
        d = eng.triangulate_points_device(mode, xy, flags | T.ALLOW_TOO_FEW, want=("xyz_f32", "mask"))
        st, bad = eng.device_status()
        h = eng.triangulate_points(mode, host, flags | T.ALLOW_TOO_FEW, want=("xyz_f32", "mask"))
        assert np.array_equal(h["xyz_f32"], d["xyz_f32"].cpu().numpy())
        assert np.array_equal(h["mask"], d["mask"].cpu().numpy().view(np.uint32))
        assert h["first_bad_frame"] == bad
    full = eng.triangulate_points(T.MATRIX, host, T.ALLOW_TOO_FEW, want=("xyz_f64", "err", "mask"))
    dd = eng.triangulate_points_device(T.MATRIX, xy, T.ALLOW_TOO_FEW, want=("xyz_f64", "err"))
    eng.device_status()
    assert np.array_equal(full["xyz_f64"], dd["xyz_f64"].cpu().numpy()) and np.array_equal(full["err"], dd["err"].cpu().numpy())


def test_unnormalised_quaternions_and_subpixel_detections(torch):
    """The reference never normalises ORIENTATION (Triangulator.cpp:15-25, Camera.h:274-287): with |q| != 1 the
    rays are weighted by |q|^4 and the extrinsic R is not orthonormal.  Sub-pixel (non-integer, non-float32)
    detections go through the double2 pixel format."""
    base = S.ring_rig(6)
    scales = [1.0, 1.3, 0.8, 1.05, 0.95, 1.2]
    cams = [T.Camera(c.cam_id, c.width, c.height, c.focal, c.position, tuple(s * v for v in c.quat)) for c, s in zip(base, scales)]
    eng = T.Engine(cams, 0)
    oc = ocams(cams)
    n = 6001
    rng = np.random.default_rng(7)
    # detections: project points with the reference's own (now non-orthonormal) P, add sub-pixel noise
    X = np.column_stack([rng.uniform(-1500, 1500, n), rng.uniform(-1500, 1500, n), rng.uniform(300, 2000, n), np.ones(n)])
    xy = np.zeros((6, n, 2))
    for c, cam in enumerate(cams):
        h = X @ cam.P.T
        xy[c] = h[:, :2] / h[:, 2:3] + rng.normal(0, 0.7, (n, 2))
    xy[rng.random((6, n)) < 0.25] = -1.0
    for omode, mode in ((O.MATRIX, T.MATRIX), (O.RAY, T.RAY)):
        ref = O.triangulate_points(oc, xy, omode, allow_too_few=True, nthreads=8, want_iters=True)
        out = eng.triangulate_points(mode, xy, T.ALLOW_TOO_FEW, want=("xyz_f64", "mask", "err"))
        assert np.array_equal(out["mask"], ref["mask"])
        ok = np.array([bin(int(m)).count("1") >= 2 for m in ref["mask"]])
        if mode == T.RAY:
            ok &= (ref["iters"] < 1000)  # where the reference's own LM converged
            assert ok.mean() > 0.7
            assert np.abs(out["xyz_f64"][ok] - ref["xyz"][ok]).max() < 1e-2  # the reference stops its LM within ~1e-3 mm of the minimiser
            np.testing.assert_allclose(out["err"][ok], ref["err"][ok], rtol=1e-5)
            ex = eng.triangulate_points(T.RAY, xy[:, :1500], T.ALLOW_TOO_FEW | T.RAY_REFERENCE_LM, want=("xyz_f64", "iters"))
            assert np.array_equal(ex["iters"], ref["iters"][:1500]) and np.array_equal(ex["xyz_f64"], ref["xyz"][:1500])
        else:
            assert rel_err(out["xyz_f64"][ok], ref["xyz"][ok]) < 1e-8
            np.testing.assert_allclose(out["err"][ok], ref["err"][ok], rtol=1e-8, atol=1e-6)
    # the float2 path rounds the sub-pixel values to float32 first: same masks, points within the FP32 pixel rounding
    o32 = eng.triangulate_points(T.MATRIX, xy.astype(np.float32), T.ALLOW_TOO_FEW, want=("xyz_f64",))
    ref = O.triangulate_points(oc, xy, O.MATRIX, allow_too_few=True, nthreads=8)
    assert np.abs(o32["xyz_f64"] - ref["xyz"]).max() < 0.05


def test_fused_gather_into_peer_memory(syn, torch):
    """The device entry point stores with plain stores, so its output may live on ANOTHER GPU: a kernel on GPU 1
    writes its shard of the points straight into GPU 0's result array over NVLink (the fused gather of DESIGN 6).
    Needs two GPUs in one box."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cams, eng0, xy, host = syn
    n = 100_000
    eng1 = T.Engine(cams, 1)
    eng1.enable_peer_access(0)
    whole = eng0.triangulate_points_device(T.MATRIX, xy[:, :2 * n].contiguous(), T.ALLOW_TOO_FEW, want=("xyz_f32",))
    eng0.device_status()
    root = torch.zeros((2 * n, 3), dtype=torch.float32, device="cuda:0")
    eng0.triangulate_points_device(T.MATRIX, xy[:, :n].contiguous(), T.ALLOW_TOO_FEW, out={"xyz_f32": root[:n]})
    xy1 = xy[:, n:2 * n].contiguous().to("cuda:1")
    with torch.cuda.device(1):
        eng1.triangulate_points_device(T.MATRIX, xy1, T.ALLOW_TOO_FEW, out={"xyz_f32": root[n:]})
        torch.cuda.synchronize()
    torch.cuda.synchronize()
    assert torch.equal(root, whole["xyz_f32"])
