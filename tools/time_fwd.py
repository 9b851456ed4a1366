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
for interp in ("bilinear", "bicubic"):
    R = RawAdvection(geo, B, V, interp, True, "fast", 6.0)
    for _ in range(3): R.forward(f, u, v, S.DT_DEFAULT)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10): R.forward(f, u, v, S.DT_DEFAULT)
    e1.record(); torch.cuda.synchronize()
    print(os.environ.get("PARADIS_SL_FWD_STRIDED", "default"), interp, "fwd ms", e0.elapsed_time(e1) / 10)
