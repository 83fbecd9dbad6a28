"""A/B of kernel shapes (TRI_VARIANT of the TUNING build, `make -C 3d-reconstruction-triangulation_b200/csrc tuning`) on
device-resident synthetic frames: CUDA-event time per launch and the largest deviation of the float3 points from the
default kernel's FP64 points.
  TRI_B200_LIB=.../libtri_b200_tuning.so python tools/ab_variants.py [--frames N] [--variants 0,3,1,2] [--cams 8]
                                                                     [--mode matrix|ray] [--precision f64|f32] [--flags N]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import tri_b200 as T  # noqa: E402
from tri_b200 import synthetic as S  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=100_000_000)
ap.add_argument("--variants", default="0,3,1,2")
ap.add_argument("--cams", type=int, default=8)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--mode", default="matrix")
ap.add_argument("--precision", default="f64")
ap.add_argument("--flags", type=int, default=0, help="extra tri_flags (e.g. 8 = TRI_RAY_CLOSED_FORM)")
a = ap.parse_args()
MODE = T.MATRIX if a.mode == "matrix" else T.RAY
FL = T.ALLOW_TOO_FEW | (T.F32 if a.precision == "f32" else 0) | a.flags
cams = S.ring_rig(a.cams)
xy = S.generate_frames(cams, a.frames, device="cuda:0")
n_chk = min(a.frames, 4_000_000)
os.environ.pop("TRI_VARIANT", None)
ref = T.Engine(cams, 0).triangulate_points_device(MODE, xy, T.ALLOW_TOO_FEW | a.flags, want=("xyz_f64", "mask"), n_frames=n_chk)
out = {"xyz_f32": torch.empty((a.frames, 3), dtype=torch.float32, device="cuda:0")}
for v in a.variants.split(","):
    os.environ["TRI_VARIANT"] = v
    eng = T.Engine(cams, 0)
    eng.triangulate_points_device(MODE, xy, FL, out=out)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.reps + 1)]
    ev[0].record()
    for i in range(a.reps):
        eng.triangulate_points_device(MODE, xy, FL, out=out)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(a.reps))
    dev = (out["xyz_f32"][:n_chk].double() - ref["xyz_f64"]).abs().max().item()
    ulp = (out["xyz_f32"][:n_chk] != ref["xyz_f64"].float()).sum().item()
    print("variant %s: median %.3f ms  min %.3f ms  %.1f GB/s   max |f32 - generic f64| = %.3e mm, %d of %d floats differ from float(generic)"
          % (v, ts[len(ts) // 2], ts[0], (8 * a.cams + 12) * a.frames / 1e6 / ts[len(ts) // 2], dev, ulp, 3 * n_chk), flush=True)
