import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import paradis_model_b200 as P
from paradis_model_b200 import synthetic as S
from paradis_model_b200.ops import RawAdvection
H, W, B, V = 721, 1440, 1, 64
lat, lon = S.make_grids(H, W, True)
geo = P.SLGeometry.from_grids(lat.cuda(), lon.cuda())
f, u, v, g = [t.cuda() for t in S.white_noise_inputs(H, W, B, V)]
R = RawAdvection(geo, B, V, "bilinear", True, "fast", 6.0)
for _ in range(3): R.backward(g, f, u, v, S.DT_DEFAULT, 3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(10): R.backward(g, f, u, v, S.DT_DEFAULT, 3)
e1.record(); torch.cuda.synchronize()
print(os.environ.get("PARADIS_SL_LIB", "default"), "fused bwd ms", e0.elapsed_time(e1) / 10)
