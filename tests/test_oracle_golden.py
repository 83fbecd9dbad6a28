"""Pins the plain-C oracle (oracle/tri_oracle.c) against the golden vectors generated from the
Python/cv2 twin (real OpenCV 4.13 arithmetic; oracle/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

import oracle_py as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def r02():
    cams = O.load_cameras(G + "/R02_D1_cameras.xml")
    offs, xy, nc, nf = O.load_dets(G + "/R02_D1_dets.npz")
    return cams, offs, xy, nc, nf


@pytest.fixture(scope="module")
def s09():
    cams = O.load_cameras(G + "/S09_D6_cameras.xml")
    offs, xy, nc, nf = O.load_dets(G + "/S09_D6_dets.npz")
    return cams, offs, xy, nc, nf


def test_camera_constants_match_cv2_twin(r02, s09):
    z = np.load(G + "/golden_cameras.npz")
    for name, cams in (("R02_D1", r02[0]), ("S09_D6", s09[0])):
        P = np.array([np.array(c.P).reshape(3, 4) for c in cams])
        K = np.array([np.array(c.K).reshape(3, 3) for c in cams])
        E = np.array([np.array(c.E).reshape(3, 4) for c in cams])
        np.testing.assert_allclose(P, z[name + "_P"], rtol=1e-13, atol=1e-9)
        np.testing.assert_allclose(K, z[name + "_K"], rtol=1e-14)
        np.testing.assert_allclose(E, z[name + "_E"], rtol=1e-13, atol=1e-10)
        fov = np.array([[c.fovx, c.fovy, c.fx, c.fy, c.cx, c.cy] for c in cams])
        np.testing.assert_allclose(fov, z[name + "_fov"], rtol=1e-14)


def test_survey_known_answer_P0(r02):
    # SURVEY.md 8(c) known-answer values (re-derived here, not trusted blindly)
    P0 = np.array(r02[0][0].P).reshape(3, 4)
    ref = np.array([[120.5761646031519, 1411.9154017817252, 293.4881764431889, 6718201.0032740785],
                    [-583.8333815353412, 599.7766406184737, -872.2514470863509, 5539785.115415021],
                    [-0.6734505494619036, 0.6772773866488861, 0.2962426352899178, 5586.58340424009]])
    np.testing.assert_allclose(P0, ref, rtol=1e-12)
    assert abs(r02[0][0].fovy - 53.16908199107405) < 1e-11


def test_batch_matrix_vs_golden(r02):
    cams, offs, xy, nc, nf = r02
    pts = O.dets_to_points(offs, xy, nc, nf)
    g = np.load(G + "/golden_R02_D1_batch.npz")
    r = O.triangulate_points(cams, pts, O.MATRIX)
    assert r["status"] == O.OK
    np.testing.assert_allclose(r["xyz"], g["matrix_xyz"], rtol=1e-11, atol=1e-9)
    np.testing.assert_allclose(r["err"], g["matrix_err"], rtol=1e-10)
    assert np.all(r["mask"] == 0xF)
    np.testing.assert_allclose(r["xyz"][0], [227.1818491131604, 465.41476576134556, 80.25051767459739], rtol=1e-10)


def test_batch_ray_vs_golden_lm_trajectory(r02):
    """The LM emulation follows cv::LMSolver's trajectory: same iteration counts, same end point."""
    cams, offs, xy, nc, nf = r02
    pts = O.dets_to_points(offs, xy, nc, nf)
    g = np.load(G + "/golden_R02_D1_batch.npz")
    r = O.triangulate_points(cams, pts, O.RAY, want_iters=True)
    assert r["status"] == O.OK
    same = r["iters"] == g["ray_iters"]
    assert same.mean() > 0.99, same.mean()
    np.testing.assert_allclose(r["xyz"][same], g["ray_xyz"][same], rtol=0, atol=1e-6)
    np.testing.assert_allclose(r["err"][same], g["ray_err"][same], rtol=1e-9)
    # every frame, converged or not, lands where the twin lands to well under a micron
    assert np.abs(r["xyz"] - g["ray_xyz"]).max() < 1e-3


def test_subsets_vs_golden(r02):
    cams, offs, xy, nc, nf = r02
    pts = O.dets_to_points(offs, xy, nc, nf)
    rows = np.load(G + "/golden_R02_D1_subsets.npz")["rows"]
    n_it_same = 0
    for row in rows:
        f, mask = int(row[0]), int(row[1])
        sub = [c for c in range(nc) if mask >> c & 1]
        X, e = O.matrix_point(cams, sub, pts[sub, f])
        np.testing.assert_allclose(X, row[2:5], rtol=1e-10, atol=1e-8)
        np.testing.assert_allclose(e, row[5], rtol=1e-9)
        Xr, er, it = O.ray_point(cams, sub, pts[sub, f])
        if it == int(row[10]):
            n_it_same += 1
            np.testing.assert_allclose(Xr, row[6:9], rtol=0, atol=1e-5)
            np.testing.assert_allclose(er, row[9], rtol=1e-8)
    assert n_it_same >= 0.97 * len(rows), (n_it_same, len(rows))


def test_ray_closed_form_agrees_with_lm_on_converged(r02):
    cams, offs, xy, nc, nf = r02
    pts = O.dets_to_points(offs, xy, nc, nf)
    for f in range(0, nf, 97):
        X, e, it = O.ray_point(cams, range(nc), pts[:, f])
        Xc = O.ray_closed_form(cams, range(nc), pts[:, f])
        if it < 1000:
            assert np.abs(X - Xc).max() < 1e-3


def _check_classify(cams, mode, n_drones, dets, gname, frames):
    offs, xy, nc, nf = dets
    g = np.load(G + "/" + gname)
    fr = g["paths"].shape[1]
    o, x, _, _ = O.slice_frames(offs, xy, nc, nf, 0, fr)
    r = O.classify(cams, mode, n_drones, o, x, nc, fr)
    assert np.array_equal(r["assign"], g["assign"])
    assert np.array_equal(r["phase"], g["phase"])
    np.testing.assert_allclose(r["paths"], g["paths"], rtol=1e-9, atol=1e-5)
    return r, g


def test_classify_r02_matrix(r02):
    cams, *dets = r02
    r, g = _check_classify(cams, O.MATRIX, 1, dets, "golden_R02_D1_classify_matrix.npz", None)
    assert np.all(r["assign"] == 1) and np.all(r["phase"][0, 1:] == 1) and r["phase"][0, 0] == 2
    assert r["stats"]["solves"] == int(g["solves"]) and r["stats"]["leaves"] == int(g["leaves"])


def test_classify_r02_ray(r02):
    cams, *dets = r02
    _check_classify(cams, O.RAY, 1, dets, "golden_R02_D1_classify_ray.npz", None)


def test_classify_s09_matrix(s09):
    cams, *dets = s09
    r, g = _check_classify(cams, O.MATRIX, 6, dets, "golden_S09_D6_classify_matrix.npz", None)
    # frame-0 picks in priority order -> paths 0..5
    want = [(4, 1, 4, 1, 1, 5, 1, 1), (2, 4, 3, 2, 3, 2, 3, 2), (5, 2, 5, 6, 0, 3, 4, 5), (6, 5, 2, 5, 2, 6, 0, 3),
            (3, 3, 1, 4, 0, 4, 0, 4), (1, 0, 6, 3, 0, 1, 2, 6)]
    assert [tuple(int(v) for v in r["assign"][p, 0]) for p in range(6)] == want
    assert r["stats"]["nodes"] == int(g["nodes"]) and r["stats"]["leaves"] == int(g["leaves"])
    assert r["stats"]["ties"] == 0 == int(g["ties"])


def test_batch_api_errors(r02):
    cams, offs, xy, nc, nf = r02
    pts = O.dets_to_points(offs, xy, nc, nf)[:, :8].copy()
    pts[1:, 3] = -1  # a single view left in frame 3
    assert O.triangulate_points(cams, pts, O.MATRIX)["status"] == O.ERR_TOO_FEW
    r = O.triangulate_points(cams, pts, O.MATRIX, allow_too_few=True)
    assert r["status"] == O.OK and r["mask"][3] == 1 and np.all(r["xyz"][3] == 0)


@pytest.mark.parametrize("name,mode,gname", [("R04_D2", O.MATRIX, "golden_R04_D2_classify_matrix.npz"),
                                             ("S01_D2_A", O.MATRIX, "golden_S01_D2_A_classify_matrix.npz"),
                                             ("R04_D2", O.RAY, "golden_R04_D2_classify_ray.npz")])
def test_classify_two_drone_datasets(name, mode, gname):
    """Two-drone sequences (real 4-camera rig and simulated 8-camera rig): the C oracle follows the cv2 twin."""
    cams = O.load_cameras(G + "/%s_cameras.xml" % name)
    dets = O.load_dets(G + "/%s_dets.npz" % name)
    r, g = _check_classify(cams, mode, 2, dets, gname, None)
    assert (r["phase"] > 0).sum() > 0.5 * r["phase"].size
