"""CPU-side checks of the product's host logic and of the C-ABI library (no compute calls: there is no
GPU here and the engine has no CPU path)."""
import ctypes
import os
import re

import numpy as np
import pytest

import tri_b200 as T
from tri_b200 import synthetic as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "tri_b200.h")).read()
    declared = sorted(set(re.findall(r"^(?:int|void|const char\*|int64_t)\s+(tri_\w+)\(", hdr, re.M)))
    assert len(declared) >= 20
    L = ctypes.CDLL(T.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert sorted(T.EXPORTS) == declared


def test_python_camera_matches_cv2_twin_golden():
    z = np.load(G + "/golden_cameras.npz")
    for name in ("R02_D1", "S09_D6"):
        cams = T.load_cameras_xml(G + "/%s_cameras.xml" % name)
        np.testing.assert_allclose(np.array([c.P for c in cams]), z[name + "_P"], rtol=1e-13, atol=1e-9)
        np.testing.assert_allclose(np.array([c.K for c in cams]), z[name + "_K"], rtol=1e-14)
        fov = np.array([[c.fovx, c.fovy, c.fx, c.fy, c.cx, c.cy] for c in cams])
        np.testing.assert_allclose(fov, z[name + "_fov"], rtol=1e-14)
    # fx is recomputed through fovx with the truncated 57.2958 (Camera.h:91,114)
    assert abs(cams[0].fx - cams[0].focal) > 1e-5


def test_camera_rejects_unset_fields():
    with pytest.raises(RuntimeError):
        T.Camera(1, 0, 1080, 1000.0, (0, 0, 0), (1, 0, 0, 0))
    with pytest.raises(RuntimeError):
        T.Camera(1, 1920, 1080, 0.0, (0, 0, 0), (1, 0, 0, 0))


@pytest.mark.skipif(T.lib().tri_device_count() > 0, reason="a GPU is present")
def test_no_device_fails_loudly():
    cams = S.ring_rig(4)
    with pytest.raises(T.TriError) as ei:
        T.Engine(cams, 0)
    assert ei.value.status == T.ERR_NO_DEVICE and "no CPU path" in str(ei.value)


def test_create_argument_checks():
    h = ctypes.c_void_p()
    assert T.lib().tri_create(0, None, 0, ctypes.byref(h)) == T.ERR_ARG
    assert T.lib().tri_create(T.MAX_CAMS + 1, None, 0, ctypes.byref(h)) == T.ERR_ARG


def test_sharded_classifier_entry_points_reject_missing_engines():
    """tri_classify_begin / _finish / _multi fail with TRI_ERR_ARG (and a message) before touching a device."""
    L = T.lib()
    assert L.tri_classify_state_bytes() == 16 * 3 * 3 * 8 + 16 * 4  # TRI_MAX_DRONES x (3-point tail) + counters
    assert L.tri_classify_begin(None, T.MATRIX, 0, 1, None, None, 0) == T.ERR_ARG
    assert L.tri_classify_finish(None, None, None, None, None, None, None) == T.ERR_ARG
    assert L.tri_classify_multi(None, 0, T.MATRIX, 0, 1, None, None, 0, None, None, None, None) == T.ERR_ARG
    assert b"engine" in L.tri_last_error()


def test_synthetic_frames_are_keyed_by_global_frame_index():
    cams = S.ring_rig(8)
    whole = S.generate_frames(cams, 3000)
    part = S.generate_frames(cams, 1000, frame0=1500, chunk=333)
    assert (whole[:, 1500:2500] == part).all()
    miss = (whole[:, :, 0] == -1)
    assert 0.17 < miss.float().mean() < 0.23
    assert ((whole[:, :, 0] == -1) == (whole[:, :, 1] == -1)).all()
    v = whole[~miss.unsqueeze(-1).expand_as(whole)]
    assert (v == v.round()).all() and v.min() >= 0 and v.max() < 1920


def test_synthetic_rig_looks_at_the_volume():
    cams = S.ring_rig(32, rings=((6000.0, 3000.0), (9000.0, 5000.0)))
    assert len(cams) == 32
    X = np.array([0.0, 0.0, 1000.0, 1.0])
    for c in cams:
        h = c.P @ X
        assert h[2] > 0 and abs(h[0] / h[2] - c.cx) < 1 and abs(h[1] / h[2] - c.cy) < 1


def test_camera_tooling_reproduces_the_fixture_xmls(tmp_path):
    """XCP / stationary_camera_data.csv -> cameras.xml: byte-identical to the committed
    fixtures the goldens were generated from; numpy-free arithmetic vs the numpy recipe of the oracle tools."""
    from tri_b200 import camera_tools as CT
    src = os.path.join(G, "camera_sources")
    CT.write_cameras_xml(CT.cameras_from_xcp(src + "/R02_D1_excerpt.xcp"), str(tmp_path / "a.xml"))
    assert open(tmp_path / "a.xml").read() == open(G + "/R02_D1_cameras.xml").read()
    CT.write_cameras_xml(CT.cameras_from_stationary_csv(src + "/S09_D6_stationary_camera_data.csv"), str(tmp_path / "b.xml"))
    got = T.load_cameras_xml(str(tmp_path / "b.xml"))
    want = T.load_cameras_xml(G + "/S09_D6_cameras.xml")
    assert len(got) == len(want) == 8
    for a, b in zip(got, want):
        assert (a.cam_id, a.width, a.height) == (b.cam_id, b.width, b.height)
        np.testing.assert_allclose(a.P, b.P, rtol=1e-12, atol=1e-7)
    assert CT.main(["camera_tools", "xcp", src + "/R02_D1_excerpt.xcp", str(tmp_path / "c.xml")]) == 0


def test_multi_drone_generator_is_keyed_by_the_global_frame_index():
    """synthetic.generate_multi_drone (BASELINE config 5's classifier input): any cut of the frame range sees the same
    detections, drones stay >= 500 mm apart and move well inside MAX_STEP, every (camera, frame) holds <= 6 detections."""
    import oracle_py as O  # slice_frames only
    cams = S.ring_rig(8)
    offs, xy, truth = S.generate_multi_drone(cams, 300, 6)
    o = offs.reshape(8, 301)
    cnt = o[:, 1:] - o[:, :-1]
    assert cnt.min() >= 0 and cnt.max() <= 6 and 4.3 < cnt.mean() < 5.2  # p_drop = 0.2
    o2, x2, t2 = S.generate_multi_drone(cams, 120, 6, frame0=180)
    a = O.slice_frames(offs, xy, 8, 300, 180, 300)
    assert np.array_equal(a[0], o2) and np.array_equal(a[1], x2) and np.array_equal(truth[:, 180:], t2)
    sep = np.linalg.norm(truth[:, None] - truth[None], axis=3)
    sep[np.arange(6), np.arange(6)] = np.inf
    assert sep.min() >= 500.0
    assert np.linalg.norm(truth[:, 1:] - truth[:, :-1], axis=2).max() < 100.0
    assert np.all(xy == np.round(xy)) and xy.min() >= 0
