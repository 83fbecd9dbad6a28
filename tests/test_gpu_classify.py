"""GPU parity of the batched DroneClassifier (tri_classify) against the oracle and the golden vectors
(generated with the cv2 twin).  Assignment indices and phases must be bit-exact."""
import os

import numpy as np
import pytest

import oracle_py as O
import tri_b200 as T

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def ocams(cams):
    return [O.make_camera(c.cam_id, c.width, c.height, c.focal, c.position, c.quat) for c in cams]


@pytest.fixture(scope="module")
def r02():
    cams = T.load_cameras_xml(G + "/R02_D1_cameras.xml")
    return cams, T.Engine(cams, 0), O.load_dets(G + "/R02_D1_dets.npz")


@pytest.fixture(scope="module")
def s09():
    cams = T.load_cameras_xml(G + "/S09_D6_cameras.xml")
    return cams, T.Engine(cams, 0), O.load_dets(G + "/S09_D6_dets.npz")


def check(r, g, atol=1e-6):
    assert np.array_equal(r["assign"], g["assign"])
    assert np.array_equal(r["phase"], g["phase"])
    np.testing.assert_allclose(r["paths"], g["paths"], rtol=1e-9, atol=atol)


def test_r02_matrix_single_drone_golden(r02):
    cams, eng, (offs, xy, nc, nf) = r02
    g = np.load(G + "/golden_R02_D1_classify_matrix.npz")
    r = eng.classify(T.MATRIX, 1, offs, xy, nf)
    check(r, g)
    assert np.all(r["assign"] == 1) and r["phase"][0, 0] == 2 and np.all(r["phase"][0, 1:] == 1)
    # for one detection per camera the classifier output equals the batch API
    pts = O.dets_to_points(offs, xy, nc, nf)
    b = eng.triangulate_points(T.MATRIX, pts, want=("xyz_f64",))
    assert np.abs(b["xyz_f64"] - r["paths"][0]).max() < 1e-9


def test_r02_ray_reference_lm_golden(r02):
    cams, eng, (offs, xy, nc, nf) = r02
    g = np.load(G + "/golden_R02_D1_classify_ray.npz")
    fr = g["paths"].shape[1]
    o, x, _, _ = O.slice_frames(offs, xy, nc, nf, 0, fr)
    r = eng.classify(T.RAY, 1, o, x, fr, T.RAY_REFERENCE_LM)
    check(r, g, atol=1e-5)
    ref = O.classify(ocams(cams), O.RAY, 1, o, x, nc, fr)
    assert np.array_equal(r["paths"], ref["paths"])  # same LM trajectory: bit-identical points


def test_r02_ray_reference_lm_long_vs_oracle(r02):
    cams, eng, (offs, xy, nc, nf) = r02
    fr = 400
    o, x, _, _ = O.slice_frames(offs, xy, nc, nf, 0, fr)
    ref = O.classify(ocams(cams), O.RAY, 1, o, x, nc, fr)
    r = eng.classify(T.RAY, 1, o, x, fr, T.RAY_REFERENCE_LM)
    assert np.array_equal(r["assign"], ref["assign"]) and np.array_equal(r["phase"], ref["phase"])
    assert np.array_equal(r["paths"], ref["paths"])


def test_s09_matrix_six_drones_golden_and_oracle(s09):
    cams, eng, (offs, xy, nc, nf) = s09
    g = np.load(G + "/golden_S09_D6_classify_matrix.npz")
    fr = g["paths"].shape[1]
    o, x, _, _ = O.slice_frames(offs, xy, nc, nf, 0, fr)
    r = eng.classify(T.MATRIX, 6, o, x, fr)
    check(r, g, atol=1e-5)
    want = [(4, 1, 4, 1, 1, 5, 1, 1), (2, 4, 3, 2, 3, 2, 3, 2), (5, 2, 5, 6, 0, 3, 4, 5), (6, 5, 2, 5, 2, 6, 0, 3),
            (3, 3, 1, 4, 0, 4, 0, 4), (1, 0, 6, 3, 0, 1, 2, 6)]
    assert [tuple(int(v) for v in r["assign"][p, 0]) for p in range(6)] == want
    assert r["stats"]["ties"] == 0


def test_s09_matrix_full_sequence_vs_oracle(s09):
    """All 3000 frames x 6 drones: indices, phases bit-exact; enumeration statistics identical."""
    cams, eng, (offs, xy, nc, nf) = s09
    ref = O.classify(ocams(cams), O.MATRIX, 6, offs, xy, nc, nf)
    r = eng.classify(T.MATRIX, 6, offs, xy, nf)
    assert np.array_equal(r["assign"], ref["assign"])
    assert np.array_equal(r["phase"], ref["phase"])
    np.testing.assert_allclose(r["paths"], ref["paths"], rtol=1e-9, atol=1e-5)
    assert r["stats"]["phase1"] == ref["stats"]["phase1"] and r["stats"]["phase2"] == ref["stats"]["phase2"]
    empties = [int((r["phase"][p] == 0).sum()) for p in range(6)]
    assert empties == [int((ref["phase"][p] == 0).sum()) for p in range(6)]


def test_full_frame_enumeration_statistics(s09):
    cams, eng, (offs, xy, nc, nf) = s09
    oc = ocams(cams)
    tot = dict(nodes=0, solves=0, leaves=0)
    frames = [0, 1, 7, 333, 1500, 2999]
    for f in frames:
        e = O.enumerate_frame(oc, O.MATRIX, offs, xy, nc, nf, f)
        for k in tot:
            tot[k] += e["stats"][k]
    got = dict(nodes=0, solves=0, leaves=0)
    for f in frames:
        o, x, _, _ = O.slice_frames(offs, xy, nc, nf, f, f + 1)
        r = eng.classify(T.MATRIX, 6, o, x, 1)
        for k in got:
            got[k] += r["stats"][k]
    assert got == tot


def test_classify_fast_ray_solver_runs(s09):
    """Closed-form ray solves inside the classifier (the throughput variant): well-formed output, every
    assigned point close to the matrix-mode point of the same frame/path where both assign."""
    cams, eng, (offs, xy, nc, nf) = s09
    fr = 60
    o, x, _, _ = O.slice_frames(offs, xy, nc, nf, 0, fr)
    r = eng.classify(T.RAY, 6, o, x, fr, T.RAY_CLOSED_FORM)
    m = eng.classify(T.MATRIX, 6, o, x, fr)
    both = (r["phase"] > 0) & (m["phase"] > 0) & np.all(r["assign"] == m["assign"], axis=2)
    assert both.mean() > 0.5
    assert np.abs(r["paths"][both] - m["paths"][both]).max() < 100.0


def test_classify_argument_checks(r02):
    cams, eng, (offs, xy, nc, nf) = r02
    with pytest.raises(T.TriError):
        eng.classify(T.MATRIX, 0, offs, xy, nf)
    r = eng.classify(T.MATRIX, 2, offs, xy, 0)
    assert r["paths"].shape == (2, 0, 3)
    # empty frames in the middle: the path gets (0,0,0) and phase 0 there (DroneClassifier.cpp:147-153)
    o = offs.reshape(nc, nf + 1)[:, :21].copy()
    cut = o.copy()
    for c in range(nc):  # drop the detections of frames 10..11 on every camera
        n10 = o[c, 12] - o[c, 10]
        cut[c, 11:] = o[c, 11:] - np.minimum(np.arange(1, 11), 2) * 0
    offs2, xy2, _, _ = O.slice_frames(offs, xy, nc, nf, 0, 20)
    o2 = offs2.reshape(nc, 21).copy()
    keep = np.ones(len(xy2), bool)
    for c in range(nc):
        keep[o2[c, 10]:o2[c, 12]] = False
    new = np.zeros_like(o2)
    base = 0
    for c in range(nc):
        cnt = np.diff(o2[c])
        cnt[10:12] = 0
        new[c, 0] = base
        new[c, 1:] = base + np.cumsum(cnt)
        base = new[c, -1]
    r = eng.classify(T.MATRIX, 1, new.reshape(-1), xy2[keep], 20)
    ref = O.classify(ocams(cams), O.MATRIX, 1, new.reshape(-1).astype(np.int32), xy2[keep], nc, 20)
    assert np.array_equal(r["phase"], ref["phase"]) and np.array_equal(r["assign"], ref["assign"])
    assert np.all(r["phase"][0, 10:12] == 0) and np.all(r["paths"][0, 10:12] == 0)


def test_accuracy_against_ground_truth(s09):
    """End-to-end sanity (the reference's only 'expected results' are accuracy figures): the six
    S09_D6 paths land within centimetres of the simulator's drone positions."""
    import torch
    from tri_b200 import evaluation as EV
    cams, eng, (offs, xy, nc, nf) = s09
    r = eng.classify(T.MATRIX, 6, offs, xy, nf)
    truth = torch.tensor(np.load(G + "/S09_D6_truth.npz")["pos_m"], dtype=torch.float64, device="cuda:0") * 1000.0
    stats = EV.evaluate(torch.tensor(r["paths"], device="cuda:0"), truth)
    assert len(stats) == 6 and all(s is not None for s in stats)
    assert sorted(s["label"] for s in stats) == list(range(6))  # every drone is followed by exactly one path
    for s in stats:
        assert s["median"] < 80.0 and s["frames_with_point"] > 0.9 * nf, s


@pytest.mark.parametrize("name,mode,flags,gname", [("R04_D2", T.MATRIX, 0, "golden_R04_D2_classify_matrix.npz"),
                                                   ("S01_D2_A", T.MATRIX, 0, "golden_S01_D2_A_classify_matrix.npz"),
                                                   ("R04_D2", T.RAY, T.RAY_REFERENCE_LM, "golden_R04_D2_classify_ray.npz")])
def test_two_drone_datasets_golden(name, mode, flags, gname):
    cams = T.load_cameras_xml(G + "/%s_cameras.xml" % name)
    eng = T.Engine(cams, 0)
    offs, xy, nc, nf = O.load_dets(G + "/%s_dets.npz" % name)
    g = np.load(G + "/" + gname)
    fr = g["paths"].shape[1]
    o, x, _, _ = O.slice_frames(offs, xy, nc, nf, 0, fr)
    r = eng.classify(mode, 2, o, x, fr, flags)
    check(r, g, atol=1e-5)


@pytest.mark.parametrize("name", ["R04_D2", "S01_D2_A"])
def test_two_drone_datasets_full_vs_oracle(name):
    cams = T.load_cameras_xml(G + "/%s_cameras.xml" % name)
    eng = T.Engine(cams, 0)
    offs, xy, nc, nf = O.load_dets(G + "/%s_dets.npz" % name)
    ref = O.classify(ocams(cams), O.MATRIX, 2, offs, xy, nc, nf)
    r = eng.classify(T.MATRIX, 2, offs, xy, nf)
    assert np.array_equal(r["assign"], ref["assign"]) and np.array_equal(r["phase"], ref["phase"])
    np.testing.assert_allclose(r["paths"], ref["paths"], rtol=1e-9, atol=1e-5)
    # more drones asked for than present: the extra paths stay empty or pick up spurious combinations identically
    o, x, _, _ = O.slice_frames(offs, xy, nc, nf, 0, min(nf, 300))
    ref3 = O.classify(ocams(cams), O.MATRIX, 3, o, x, nc, min(nf, 300))
    r3 = eng.classify(T.MATRIX, 3, o, x, min(nf, 300))
    assert np.array_equal(r3["assign"], ref3["assign"]) and np.array_equal(r3["phase"], ref3["phase"])


def test_s09_ray_reference_lm_matches_oracle_short(s09):
    """--triangulator ray with 6 drones and 8 cameras: thousands of cv::LMSolver emulations per frame, many of
    them the non-converging 2-view kind -- decisions and points identical to the oracle."""
    cams, eng, (offs, xy, nc, nf) = s09
    fr = 6
    o, x, _, _ = O.slice_frames(offs, xy, nc, nf, 0, fr)
    ref = O.classify(ocams(cams), O.RAY, 6, o, x, nc, fr)
    r = eng.classify(T.RAY, 6, o, x, fr, T.RAY_REFERENCE_LM)
    assert np.array_equal(r["assign"], ref["assign"]) and np.array_equal(r["phase"], ref["phase"])
    assert np.array_equal(r["paths"], ref["paths"])
    assert r["stats"]["lm_iters"] > 1e5


def test_long_sequence_crosses_frame_batches(s09):
    """More frames than one enumeration batch (8192): the tracking state (last point + tail per path) is
    carried across batches on the device.  S09_D6 played forwards, backwards, forwards, backwards."""
    cams, eng, (offs, xy, nc, nf) = s09
    o = offs.reshape(nc, nf + 1)
    reps = 4
    new_offs = np.zeros((nc, reps * nf + 1), np.int64)
    parts = []
    base = 0
    for c in range(nc):
        cnt = np.diff(o[c])
        seq_cnt, seq_xy = [], []
        for r in range(reps):
            order = range(nf) if r % 2 == 0 else range(nf - 1, -1, -1)
            for f in order:
                seq_cnt.append(cnt[f])
                seq_xy.append(xy[o[c, f]:o[c, f + 1]])
        new_offs[c, 0] = base
        new_offs[c, 1:] = base + np.cumsum(seq_cnt)
        base = new_offs[c, -1]
        parts.append(np.concatenate(seq_xy, axis=0))
    big_offs = new_offs.astype(np.int32).reshape(-1)
    big_xy = np.concatenate(parts, axis=0)
    n = reps * nf
    assert n > 8192
    ref = O.classify(ocams(cams), O.MATRIX, 6, big_offs, big_xy, nc, n)
    r = eng.classify(T.MATRIX, 6, big_offs, big_xy, n)
    assert np.array_equal(r["assign"], ref["assign"]) and np.array_equal(r["phase"], ref["phase"])
    np.testing.assert_allclose(r["paths"], ref["paths"], rtol=1e-9, atol=1e-5)


@pytest.mark.parametrize("world", [2, 3, 5])
def test_frame_sharded_chain_equals_whole_sequence(s09, world):
    """every shard's candidates are enumerated independently (tri_classify_begin), the shards are linked
    in order with only the tracking state handed along (tri_classify_finish) -- bit for bit the unsharded result.
    One engine per shard; on a one-GPU box they share the device, with more GPUs they are spread over them."""
    import torch
    from tri_b200 import sharding as SH
    cams, eng, (offs, xy, nc, nf) = s09
    whole = eng.classify(T.MATRIX, 6, offs, xy, nf)
    n_dev = torch.cuda.device_count()
    engines = [T.Engine(cams, g % n_dev) for g in range(world)]
    r = SH.classify_chain(engines, T.MATRIX, 6, offs, xy, nf)
    assert np.array_equal(r["assign"], whole["assign"]) and np.array_equal(r["phase"], whole["phase"])
    assert np.array_equal(r["paths"], whole["paths"])
    for k in ("nodes", "solves", "leaves", "phase1", "phase2", "ties"):
        assert r["stats"][k] == whole["stats"][k], k
    # the C entry point that does the same from one process (a host thread per engine for the enumeration)
    m = T.classify_multi(engines, T.MATRIX, 6, offs, xy, nf)
    assert np.array_equal(m["assign"], whole["assign"]) and np.array_equal(m["phase"], whole["phase"]) and np.array_equal(m["paths"], whole["paths"])
    assert m["stats"]["leaves"] == whole["stats"]["leaves"] and m["stats"]["phase1"] == whole["stats"]["phase1"]
    # the fast ray solver and a single-drone sequence through the same chain
    w2 = eng.classify(T.RAY, 6, offs, xy, nf, T.RAY_CLOSED_FORM)
    r2 = SH.classify_chain(engines, T.RAY, 6, offs, xy, nf, T.RAY_CLOSED_FORM)
    assert np.array_equal(r2["assign"], w2["assign"]) and np.array_equal(r2["paths"], w2["paths"])


def test_frame_sharded_edge_cases(r02):
    """Shards shorter than the tail, an empty shard, finish without begin."""
    from tri_b200 import sharding as SH
    cams, eng, (offs, xy, nc, nf) = r02
    o, x, _, n = O.slice_frames(offs, xy, nc, nf, 0, 7)
    whole = eng.classify(T.MATRIX, 1, o, x, n)
    engines = [T.Engine(cams, 0) for _ in range(9)]  # 9 shards over 7 frames: some are empty
    r = SH.classify_chain(engines, T.MATRIX, 1, o, x, n)
    assert np.array_equal(r["assign"], whole["assign"]) and np.array_equal(r["paths"], whole["paths"])
    fresh = T.Engine(cams, 0)
    fresh._cls_job = (T.MATRIX, 1, 7)
    with pytest.raises(T.TriError):
        fresh.classify_finish(None)  # tri_classify_finish without tri_classify_begin


def test_frame_sharded_two_processes_nccl(s09, tmp_path):
    """One process per GPU, the tracking state sent rank to rank over NCCL (needs 2 GPUs)."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29533", os.path.join(root, "tests", "_dist_classify_worker.py")],
                         capture_output=True, text=True, timeout=600)
    assert "DIST_CLASSIFY_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_synthetic_six_drones_and_independent_sequences():
    """BASELINE config 5's classifier input (synthetic.generate_multi_drone: 6 drones, permuted detections, 20 % dropped):
    tri_classify equals the oracle on it, and tri_classify_sequences over several recordings laid back to back equals
    tri_classify on each of them alone."""
    from tri_b200 import synthetic as S
    cams = S.ring_rig(8)
    nf = 360
    offs, xy, truth = S.generate_multi_drone(cams, nf, 6)
    eng = T.Engine(cams, 0)
    whole = eng.classify(T.MATRIX, 6, offs, xy, nf)
    ref = O.classify(ocams(cams), O.MATRIX, 6, offs, xy, 8, nf)
    assert np.array_equal(whole["assign"], ref["assign"]) and np.array_equal(whole["phase"], ref["phase"])
    np.testing.assert_allclose(whole["paths"], ref["paths"], rtol=1e-9, atol=1e-6)
    # every drone is tracked to a few millimetres of the simulated flight
    for p in range(6):
        d = np.linalg.norm(whole["paths"][p][:, None, :] - truth.transpose(1, 0, 2), axis=2)
        assert np.median(d.min(axis=1)) < 20.0
    bounds = [0, 100, 100, 101, 250, nf]  # an empty and a one-frame recording among them
    many = eng.classify_sequences(T.MATRIX, 6, bounds, offs, xy, nf)
    for a, b in zip(bounds[:-1], bounds[1:]):
        if b == a:
            continue
        o, x, _, n = O.slice_frames(offs, xy, 8, nf, a, b)
        one = eng.classify(T.MATRIX, 6, o, x, n)
        assert np.array_equal(many["assign"][:, a:b], one["assign"]) and np.array_equal(many["phase"][:, a:b], one["phase"])
        assert np.array_equal(many["paths"][:, a:b], one["paths"])


@pytest.mark.parametrize("mode,flags,frames", [(T.MATRIX, 0, 400), (T.RAY, T.RAY_CLOSED_FORM, 120), (T.RAY, 0, 25)])
def test_lazy_search_equals_the_enumeration_at_8_cameras(s09, mode, flags, frames):
    """TRI_CLS_LAZY (the classifier of the 17..32-camera rigs: branch and bound over the reference's tree instead of
    enumerating it) returns bit-identical paths, assignments and phases where the enumerating classifier can run."""
    cams, eng, (offs, xy, nc, nf) = s09
    o, x, _, _ = O.slice_frames(offs, xy, nc, nf, 0, frames)
    a = eng.classify(mode, 6, o, x, frames, flags)
    b = eng.classify(mode, 6, o, x, frames, flags | T.CLS_LAZY)
    assert np.array_equal(a["assign"], b["assign"]) and np.array_equal(a["phase"], b["phase"])
    assert np.array_equal(a["paths"], b["paths"])
    assert (a["stats"]["phase1"], a["stats"]["phase2"]) == (b["stats"]["phase1"], b["stats"]["phase2"])


def test_lazy_search_on_12_and_32_camera_rigs():
    """12 cameras: the lazy search against the enumerating classifier (still runnable there: <= 16 cameras) on six simulated
    drones, bit for bit, and both against the oracle's decisions on a prefix.  32 cameras (BASELINE config 5): only the lazy
    search can run -- the enumeration the reference does is ~6 * 2^26 leaves per frame -- so it is checked against the
    simulated flight: every drone tracked in every frame, a few millimetres from the truth, with ~25 cameras per point."""
    from tri_b200 import synthetic as S
    cams = S.ring_rig(12)
    nf = 60
    offs, xy, truth = S.generate_multi_drone(cams, nf, 6)
    eng = T.Engine(cams, 0)
    a = eng.classify(T.MATRIX, 6, offs, xy, nf)
    b = eng.classify(T.MATRIX, 6, offs, xy, nf, T.CLS_LAZY)
    assert np.array_equal(a["assign"], b["assign"]) and np.array_equal(a["phase"], b["phase"]) and np.array_equal(a["paths"], b["paths"])
    o, x, _, _ = O.slice_frames(offs, xy, 12, nf, 0, 6)
    ref = O.classify(ocams(cams), O.MATRIX, 6, o, x, 12, 6)
    assert np.array_equal(ref["assign"], b["assign"][:, :6]) and np.array_equal(ref["phase"], b["phase"][:, :6])
    np.testing.assert_allclose(b["paths"][:, :6], ref["paths"], rtol=1e-9, atol=1e-6)
    # a sequence cut into recordings goes through the same search
    bounds = [0, 25, 25, 60]
    m = eng.classify_sequences(T.MATRIX, 6, bounds, offs, xy, nf, T.CLS_LAZY)
    for f0, f1 in ((0, 25), (25, 60)):
        oo, xx, _, n = O.slice_frames(offs, xy, 12, nf, f0, f1)
        one = eng.classify(T.MATRIX, 6, oo, xx, n, T.CLS_LAZY)
        assert np.array_equal(m["assign"][:, f0:f1], one["assign"]) and np.array_equal(m["paths"][:, f0:f1], one["paths"])

    cams32 = S.ring_rig(32, rings=((6000.0, 3000.0), (9000.0, 5000.0)))
    nf = 150
    offs, xy, truth = S.generate_multi_drone(cams32, nf, 6)
    eng32 = T.Engine(cams32, 0)
    for mode, flags in ((T.MATRIX, 0), (T.RAY, T.RAY_CLOSED_FORM)):
        r = eng32.classify(mode, 6, offs, xy, nf, flags)
        assert np.all(r["phase"][:, 0] == 2) and np.all(r["phase"][:, 1:] == 1)  # re-initialised once, tracked ever after
        for p in range(6):
            d = np.linalg.norm(r["paths"][p][:, None, :] - truth.transpose(1, 0, 2), axis=2)
            assert np.median(d.min(axis=1)) < 10.0 and d.min(axis=1)[1:].max() < 60.0 and d.min(axis=1)[0] < 200.0  # frame 0 is the greedy re-initialisation
        used = (r["assign"] > 0).sum(axis=2)
        assert used.mean() > 20  # ~80 % of 32 cameras per point
        # every assigned detection exists, and no detection serves two drones in a frame
        o2 = offs.reshape(32, nf + 1)
        cnt = (o2[:, 1:] - o2[:, :-1]).T  # [frame][cam]
        assert np.all(r["assign"] <= cnt[None, :, :])
        for f in range(0, nf, 7):
            for c in range(32):
                ks = [int(k) for k in r["assign"][:, f, c] if k > 0]
                assert len(ks) == len(set(ks))

def test_twelve_drones_more_paths_than_half_the_warps():
    """Twelve drones on a 4-camera rig: more tracked paths than the linking kernel has warps per path to spare (one warp each, no
    second one to share a scan), twelve detections per camera and frame; enumeration + linking and the lazy search against the
    oracle, bit for bit, and a run with every other frame empty (paths re-initialise, tails restart)."""
    from tri_b200 import synthetic as S
    cams = S.ring_rig(4)
    nf = 40
    offs, xy, truth = S.generate_multi_drone(cams, nf, 12)
    eng = T.Engine(cams, 0)
    ref = O.classify(ocams(cams), O.MATRIX, 12, offs, xy, 4, nf)
    for flags in (0, T.CLS_LAZY):
        r = eng.classify(T.MATRIX, 12, offs, xy, nf, flags)
        assert np.array_equal(ref["assign"], r["assign"]) and np.array_equal(ref["phase"], r["phase"])
        np.testing.assert_allclose(r["paths"], ref["paths"], rtol=1e-9, atol=1e-6)
    assert int((ref["phase"] == 1).sum()) > 6 * nf  # most of the twelve are tracked most of the time
    # every other frame without a single detection
    o2 = np.asarray(offs, np.int64).reshape(4, nf + 1).copy()
    keep = np.ones(len(xy), bool)
    for c in range(4):
        cnt = np.diff(o2[c])
        for f in range(1, nf, 2):
            keep[o2[c, f]:o2[c, f + 1]] = False
            cnt[f] = 0
        o2[c, 1:] = o2[c, 0] + np.cumsum(cnt)
    base = 0
    for c in range(4):  # cameras back to back again
        n_c = o2[c, -1] - o2[c, 0]
        o2[c] = o2[c] - o2[c, 0] + base
        base += n_c
    xy2 = np.asarray(xy)[keep]
    o2 = o2.astype(np.int32).reshape(-1)
    ref2 = O.classify(ocams(cams), O.MATRIX, 12, o2, xy2, 4, nf)
    r2 = eng.classify(T.MATRIX, 12, o2, xy2, nf)
    assert np.array_equal(ref2["assign"], r2["assign"]) and np.array_equal(ref2["phase"], r2["phase"])
