"""One forward+backward of the hot path for ncu (GPU box).  usage: prof_step.py [V] [interp] [math] [steps] [cfl]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import paradis_model_b200 as P
from paradis_model_b200 import synthetic as S
from paradis_model_b200.ops import RawAdvection
V = int(sys.argv[1]) if len(sys.argv) > 1 else 16
interp = sys.argv[2] if len(sys.argv) > 2 else "bilinear"
math = sys.argv[3] if len(sys.argv) > 3 else "fast"
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
cfl = float(sys.argv[5]) if len(sys.argv) > 5 else 6.0
H, W, B = 721, 1440, 1
lat, lon = S.make_grids(H, W, True)
geo = P.SLGeometry.from_grids(lat.cuda(), lon.cuda())
f, u, v, g = [t.cuda() for t in S.white_noise_inputs(H, W, B, V)]
R = RawAdvection(geo, B, V, interp, True, math, cfl)
for _ in range(steps):
    R.forward(f, u, v, S.DT_DEFAULT)
    R.backward(g, f, u, v, S.DT_DEFAULT, 3)
torch.cuda.synchronize()
P.check_status()
print("done")
