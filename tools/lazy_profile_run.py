"""One 32-camera classification through the lazy search, the target of ncu -k regex:lazy_link_kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tri_b200 as T
from tri_b200 import synthetic as S
cams = S.ring_rig(32, rings=((6000.0, 3000.0), (9000.0, 5000.0)))
eng = T.Engine(cams, 0)
nf = 500
offs, xy, truth = S.generate_multi_drone(cams, nf, 6)
eng.classify(T.MATRIX, 6, offs, xy, nf)
r = eng.classify(T.MATRIX, 6, offs, xy, nf)
print(r["stats"])
