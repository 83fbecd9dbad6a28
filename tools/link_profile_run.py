"""One classification for ncu (-k regex:link_kernel --import-source on): S09_D6 (default) or --synthetic N frames of 6 drones."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py as O
import tri_b200 as T
if "--synthetic" in sys.argv:
    from tri_b200 import synthetic as S
    nf = int(sys.argv[sys.argv.index("--synthetic") + 1])
    cams = S.ring_rig(8)
    offs, xy, _ = S.generate_multi_drone(cams, nf, 6)
else:
    G = os.path.join(ROOT, "tests", "golden")
    cams = T.load_cameras_xml(G + "/S09_D6_cameras.xml")
    offs, xy, nc, nf = O.load_dets(G + "/S09_D6_dets.npz")
eng = T.Engine(cams, 0)
eng.classify(T.MATRIX, 6, offs, xy, nf, 0)
t0 = time.perf_counter(); eng.classify(T.MATRIX, 6, offs, xy, nf, 0); print("gpu_s", time.perf_counter() - t0)
