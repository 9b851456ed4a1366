"""Small forward+backward cases for compute-sanitizer (memcheck / racecheck / initcheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import paradis_model_b200 as P
from paradis_model_b200 import synthetic as S
from paradis_model_b200.ops import RawAdvection
for (H, W, B, V, poles, cfl, interp) in [(96, 192, 1, 2, True, 3.0, "bilinear"), (64, 160, 2, 1, False, 2.0, "bicubic"),
                                         (24, 18, 1, 2, True, 0.0, "bilinear")]:
    lat, lon = S.make_grids(H, W, poles)
    geo = P.SLGeometry.from_grids(lat.cuda(), lon.cuda())
    f, u, v, g = [t.cuda() for t in S.white_noise_inputs(H, W, B, V, cells_sigma=1.0, cells_clip=max(cfl / 1.5, 1.0))]
    R = RawAdvection(geo, B, V, interp, True, "fast", cfl)
    R.forward(f, u, v, S.DT_DEFAULT)
    R.backward(g, f, u, v, S.DT_DEFAULT, 3)
    torch.cuda.synchronize()
    P.check_status()
    x = torch.randn(1, 2, H, W, device="cuda", requires_grad=True)
    P.geocyclic_pad(x, 2).sum().backward()
    print("case", H, W, interp, "ok", float(R.gfield.abs().sum()))
