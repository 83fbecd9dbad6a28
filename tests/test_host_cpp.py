"""The C++ host mirror of the reference interface (3d-reconstruction-triangulation_b200/host):
CPU checks of the loaders (no engine calls) and, on the GPU, the CLI clone end to end."""
import glob
import os
import struct
import subprocess

import numpy as np
import pytest

import oracle_py as O
import py_twin
import tri_b200 as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "3d-reconstruction-triangulation_b200", "host")
G = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def host_bins():
    subprocess.check_call(["make", "-s", "-C", HOST])
    return os.path.join(HOST, "host_selftest"), os.path.join(HOST, "tri_main"), os.path.join(HOST, "host_gpu_selftest")


def test_cpp_loaders_match_python_host_and_twin(host_bins):
    out = subprocess.check_output([host_bins[0], G + "/S09_D6_cameras.xml", G + "/csv_sample"], text=True).splitlines()
    cams = T.load_cameras_xml(G + "/S09_D6_cameras.xml")
    rows = [l.split() for l in out if l.startswith("cam ")]
    assert out[0] == "cameras %d" % len(cams) and len(rows) == len(cams)
    for r, c in zip(rows, cams):
        assert (int(r[1]), int(r[2]), int(r[3])) == (c.cam_id, c.width, c.height)
        assert float(r[4]) == c.fovy and float(r[5]) == c.fx
        assert np.array_equal(np.array([float(v) for v in r[6:18]]), c.P.reshape(-1))  # bit-identical constants
    files = sorted(glob.glob(G + "/csv_sample/*.csv"))
    want = py_twin.read_csv_files(files)  # the twin's restatement of DetectionsContainer::readFiles
    n_cam, n_frames = (int(v) for v in [l for l in out if l.startswith("container")][0].split()[1:])
    assert n_cam == len(want) and n_frames == len(want[0])
    dets = {(int(l.split()[1]), int(l.split()[2])): [float(v) for v in l.split()[4:]] for l in out if l.startswith("det ")}
    for c in range(n_cam):
        for f in range(n_frames):
            assert dets[(c, f)] == [v for p in want[c][f] for v in p]
    assert len({len(want[c][f]) for c in range(n_cam) for f in range(n_frames)}) > 1  # ragged detection counts
    n_det = sum(len(want[c][f]) for c in range(n_cam) for f in range(n_frames))
    assert [l for l in out if l.startswith("csr")][0] == "csr %d %d %d" % (n_cam * (n_frames + 1), n_det, n_det)
    # the batch adapter's packing decision: ushort2 (TRI_PIX_U16) / float2 (0) / double2 (TRI_PIX_F64)
    assert [l for l in out if l.startswith("pixfmt")][0] == "pixfmt %d 0 0 0 %d %d" % (T.PIX_U16, T.PIX_F64, T.PIX_U16)


def test_cpp_csv_reader_gaps_and_truncation(host_bins, tmp_path):
    """Missing frame numbers become empty frames, id-only rows are frames without detections, every token
    goes through stoi (0.97 -> 0, 12.9 -> 12): DetectionsContainer.cpp:19-76."""
    a = "0,1,2,3,4,100,200,0.97\n3,1,2,3,4,101.9,201,0.5,5,6,7,8,300,400,0.99\n4\n6,1,1,1,1,7,8,1\n"
    b = "0,1,2,3,4,110,210,0.97\n1,1,2,3,4,111,211,0.9\n2\n3\n4,9,9,9,9,114,214,0.3\n5\n6,1,1,1,1,17,18,1\n"
    (tmp_path / "camA.csv").write_text(a)
    (tmp_path / "camB.csv").write_text(b)
    out = subprocess.check_output([host_bins[0], G + "/R02_D1_cameras.xml", str(tmp_path)], text=True).splitlines()
    want = py_twin.read_csv_files([str(tmp_path / "camA.csv"), str(tmp_path / "camB.csv")])
    assert [l for l in out if l.startswith("container")][0] == "container 2 7"
    dets = {(int(l.split()[1]), int(l.split()[2])): [float(v) for v in l.split()[4:]] for l in out if l.startswith("det ")}
    for c in range(2):
        for f in range(7):
            assert dets[(c, f)] == [v for p in want[c][f] for v in p]
    assert dets[(0, 1)] == [] and dets[(0, 2)] == [] and dets[(0, 4)] == [] and dets[(0, 3)] == [101.0, 201.0, 300.0, 400.0]
    # unequal frame counts are rejected like the reference does
    (tmp_path / "camB.csv").write_text(b + "7,1,1,1,1,1,1,1\n")
    r = subprocess.run([host_bins[0], G + "/R02_D1_cameras.xml", str(tmp_path)], capture_output=True, text=True)
    assert r.returncode != 0 and "Number of frames on all cameras must be the same" in r.stderr


def test_cli_usage_errors(host_bins):
    r = subprocess.run([host_bins[1]], capture_output=True, text=True)
    assert r.returncode == 1 and "cameras_path" in r.stderr


def write_csvs(tmp, dets_npz, n_frames=None):
    offs, xy, nc, nf = O.load_dets(dets_npz)
    nf = n_frames or nf
    o = offs.reshape(nc, -1)
    names = [str(n) for n in np.load(dets_npz)["files"]]
    for c in range(nc):
        with open(os.path.join(tmp, names[c]), "w") as fh:
            for f in range(nf):
                rec = [str(f)]
                for x, y in xy[o[c, f]:o[c, f + 1]]:
                    rec += [str(int(x) - 10), str(int(y) - 5), "20", "10", str(int(x)), str(int(y)), "0.9"]
                if len(rec) > 1:  # the reference's CSVs simply omit frames without detections... or hold only the id
                    fh.write(",".join(rec) + "\n")
                elif f % 2 == 0:
                    fh.write(str(f) + "\n")
    return nc, nf


def read_dump(path):
    raw = open(path, "rb").read()
    d, f, c = struct.unpack("iii", raw[:12])
    p = np.frombuffer(raw, np.float64, d * f * 3, 12).reshape(d, f, 3)
    a = np.frombuffer(raw, np.int8, d * f * c, 12 + 8 * d * f * 3).reshape(d, f, c)
    ph = np.frombuffer(raw, np.uint8, d * f, 12 + 8 * d * f * 3 + d * f * c).reshape(d, f)
    return p, a, ph


@pytest.mark.gpu
@pytest.mark.parametrize("kind,golden,frames,drones,data", [
    ("matrix", "golden_R02_D1_classify_matrix.npz", None, 1, "R02_D1"),
    ("ray", "golden_R02_D1_classify_ray.npz", 40, 1, "R02_D1"),
    ("matrix", "golden_S09_D6_classify_matrix.npz", 120, 6, "S09_D6"),
])
def test_cli_end_to_end(host_bins, tmp_path, kind, golden, frames, drones, data):
    """./main cameras.xml data/ --n_drones N --triangulator K: PLY files + 'Execution time', same paths and
    assignment indices as the golden vectors."""
    csv_dir = tmp_path / "data"
    csv_dir.mkdir()
    nc, nf = write_csvs(str(csv_dir), G + "/%s_dets.npz" % data, frames)
    r = subprocess.run([host_bins[1], G + "/%s_cameras.xml" % data, str(csv_dir), "--n_drones", str(drones), "--triangulator", kind,
                        "--dump", str(tmp_path / "dump.bin")], capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr
    assert "Execution time: " in r.stdout
    g = np.load(G + "/" + golden)
    p, a, ph = read_dump(str(tmp_path / "dump.bin"))
    assert np.array_equal(a, g["assign"]) and np.array_equal(ph, g["phase"])
    np.testing.assert_allclose(p, g["paths"], rtol=1e-9, atol=1e-5)
    for d in range(drones):
        lines = open(tmp_path / "results" / ("drone%d.ply" % (d + 1))).read().splitlines()
        assert lines[0] == "ply" and lines[2] == "element vertex %d" % nf and lines[8] == "end_header"
        got = np.array([[float(v) for v in l.split()] for l in lines[9:]])
        np.testing.assert_allclose(got, g["paths"][d], rtol=2e-5, atol=1e-3)  # default ostream precision: 6 digits


@pytest.mark.gpu
def test_cli_frame_sharded_classifier(host_bins, tmp_path):
    """TRI_B200_GPUS=3 ./main ...: the C++ DroneClassifier adapter shards the sequence over three engines
    (tri_classify_multi) -- same dump as the golden vectors."""
    csv_dir = tmp_path / "data"
    csv_dir.mkdir()
    nc, nf = write_csvs(str(csv_dir), G + "/S09_D6_dets.npz", 120)
    env = dict(os.environ, TRI_B200_GPUS="3")
    r = subprocess.run([host_bins[1], G + "/S09_D6_cameras.xml", str(csv_dir), "--n_drones", "6", "--triangulator", "matrix",
                        "--dump", str(tmp_path / "dump.bin")], capture_output=True, text=True, cwd=str(tmp_path), env=env)
    assert r.returncode == 0, r.stderr
    g = np.load(G + "/golden_S09_D6_classify_matrix.npz")
    p, a, ph = read_dump(str(tmp_path / "dump.bin"))
    assert np.array_equal(a, g["assign"]) and np.array_equal(ph, g["phase"])
    np.testing.assert_allclose(p, g["paths"], rtol=1e-9, atol=1e-5)


@pytest.mark.gpu
def test_cli_rejects_unknown_triangulator(host_bins, tmp_path):
    csv_dir = tmp_path / "data"
    csv_dir.mkdir()
    write_csvs(str(csv_dir), G + "/R02_D1_dets.npz", 5)
    r = subprocess.run([host_bins[1], G + "/R02_D1_cameras.xml", str(csv_dir), "--triangulator", "svd"], capture_output=True, text=True,
                       cwd=str(tmp_path))
    assert r.returncode != 0 and "Invalid --triangulator argument" in r.stderr


@pytest.mark.gpu
def test_cpp_triangulator_adapters(host_bins, tmp_path):
    """The C++ MatrixTriangulator / RayTriangulator adapters called the way the reference's code calls them:
    triangulatePoints, triangulatePoint on a subset, static getDistFromRay, and the reference's exception texts."""
    csv_dir = tmp_path / "data"
    csv_dir.mkdir()
    nc, nf = write_csvs(str(csv_dir), G + "/R02_D1_dets.npz", 60)
    out = subprocess.check_output([host_bins[2], G + "/R02_D1_cameras.xml", str(csv_dir)], text=True).splitlines()
    assert out[0] == "types matrix ray 4"
    g = np.load(G + "/golden_R02_D1_batch.npz")
    m = np.array([[float(v) for v in l.split()[2:]] for l in out if l.startswith("matrix ")])
    r = np.array([[float(v) for v in l.split()[2:]] for l in out if l.startswith("ray ")])
    assert m.shape == (60, 3) and r.shape == (60, 3)
    np.testing.assert_allclose(m, g["matrix_xyz"][:60], rtol=1e-9, atol=1e-7)
    same = np.ones(60, bool)
    np.testing.assert_allclose(r[same], g["ray_xyz"][:60][same], rtol=0, atol=1e-5)  # RayTriangulator defaults to the exact LM
    cams = O.load_cameras(G + "/R02_D1_cameras.xml")
    offs, xy, _, nfull = O.load_dets(G + "/R02_D1_dets.npz")
    pts = O.dets_to_points(offs, xy, nc, nfull)
    # integer detections travel as ushort2; off the integer grid the adapter packs float2 / double2
    for tag, shift in (("matrix_f32 ", 0.25), ("matrix_f64 ", 1e-7)):
        got = np.array([[float(v) for v in l.split()[2:]] for l in out if l.startswith(tag)])
        want = O.triangulate_points(cams, np.ascontiguousarray(pts[:, :60] + shift), O.MATRIX)["xyz"]
        assert got.shape == (60, 3)
        np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-7)
    Xm, em = O.matrix_point(cams, [0, 2], pts[[0, 2], 0])
    got = [float(v) for v in [l for l in out if l.startswith("point_matrix")][0].split()[1:]]
    np.testing.assert_allclose(got, [*Xm, em], rtol=1e-9)
    Xr, er, _ = O.ray_point(cams, [0, 2], pts[[0, 2], 0])
    got = [float(v) for v in [l for l in out if l.startswith("point_ray")][0].split()[1:]]
    assert got == [*Xr, er]  # trajectory-exact
    want = O.lib().orc_dist_from_ray(O.C.byref(cams[1]), O.C.c_double(pts[1, 0, 0]), O.C.c_double(pts[1, 0, 1]), O._p(np.ascontiguousarray(Xm)))
    d = float([l for l in out if l.startswith("dist ")][0].split()[1])
    assert abs(d - want) <= 1e-9 * max(1.0, abs(want))
    assert "throw_dim Every camera should have the same number of points" in out
    assert "throw_few Too few rays are found" in out and "throw_few Too few detections are found" in out
    assert "throw_one Too few rays are found" in out


def test_cli_flag_forms_follow_the_reference_parser(host_bins):
    """src/main.cpp:16-41 parses with p-ranav argparse: -h/--help and -v/--version exit 0, '--flag=value' is accepted,
    an unknown flag or a non-integer --n_drones is an error (usage on stderr, exit code 1)."""
    exe = host_bins[1]
    r = subprocess.run([exe, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "cameras_path" in r.stderr + r.stdout and "--n_drones" in r.stderr + r.stdout
    r = subprocess.run([exe, "-v"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "1.0"
    r = subprocess.run([exe, "a.xml", "data", "--bogus"], capture_output=True, text=True)
    assert r.returncode == 1 and "Unknown argument: --bogus" in r.stderr
    r = subprocess.run([exe, "a.xml", "data", "--n_drones=x3"], capture_output=True, text=True)
    assert r.returncode == 1 and "pattern" in r.stderr
    r = subprocess.run([exe, "a.xml", "--n_drones"], capture_output=True, text=True)
    assert r.returncode == 1 and "1 argument(s) expected. 0 provided." in r.stderr
    r = subprocess.run([exe, "a.xml", "b", "c"], capture_output=True, text=True)
    assert r.returncode == 1 and "Maximum number of positional arguments exceeded" in r.stderr


def test_cpp_xml_loader_skips_what_pugixml_skips(host_bins, tmp_path):
    """Comments (with camera-looking text inside), CDATA, a processing instruction, single-quoted attributes, a '>' inside
    an attribute value, entities and cameras without a ControlFrame: the scanner sees what the reference's pugixml
    loader sees (src/utils.cpp:46-92) -- checked against the reference's own loader when oracle/_ref is built."""
    src = open(G + "/R02_D1_cameras.xml").read()
    cams = T.load_cameras_xml(G + "/R02_D1_cameras.xml")
    body = src[src.index("<Cameras"):]
    first = body.index("<Camera ")
    tricky = ('<?xml version="1.0"?>\n<!-- <Camera DEVICEID="1"><ControlFrames><ControlFrame FOCAL_LENGTH="1"/></ControlFrames></Camera> -->\n'
              + body[:first]
              + '<Camera DEVICEID="777" NOTE="a > b"><ControlFrames/></Camera>\n<![CDATA[ <Camera DEVICEID="5"> ]]>\n'
              + body[first:].replace('DEVICEID="', "DEVICEID='", 1).replace('"', "'", 1))
    p = tmp_path / "tricky.xml"
    p.write_text(tricky)
    out = subprocess.check_output([host_bins[0], str(p), G + "/csv_sample"], text=True).splitlines()
    rows = [l.split() for l in out if l.startswith("cam ")]
    assert out[0] == "cameras %d" % len(cams)
    for r, c in zip(rows, cams):
        assert int(r[1]) == c.cam_id and np.array_equal(np.array([float(v) for v in r[6:18]]), c.P.reshape(-1))
    import ref_py as R
    if R.available():
        ref = R.Reference(xml=str(p), mode=R.MATRIX)
        assert ref.n_cams == len(cams)
        for i, c in enumerate(cams):
            assert np.array_equal(ref.camera(i)["P"], c.P)
