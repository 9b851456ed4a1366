import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import paradis_model_b200 as P
from paradis_model_b200 import synthetic as S
from paradis_model_b200.ops import RawAdvection
V = int(sys.argv[1]) if len(sys.argv) > 1 else 64
interp = sys.argv[2] if len(sys.argv) > 2 else "bilinear"
H, W, B = 721, 1440, 1
lat, lon = S.make_grids(H, W, True)
geo = P.SLGeometry.from_grids(lat.cuda(), lon.cuda())
f, u, v, g = [t.cuda() for t in S.white_noise_inputs(H, W, B, V)]
R = RawAdvection(geo, B, V, interp, True, "fast", 6.0)
for _ in range(2):
    R.backward(g, f, u, v, S.DT_DEFAULT, 3)
torch.cuda.synchronize()
P.check_status()
