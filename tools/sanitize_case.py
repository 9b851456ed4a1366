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
    R.backward(g, f, u, v, S.DT_DEFAULT, 3)          # the library's own choice of fused backward
    R.backward(g, f, u, v, S.DT_DEFAULT, 3 | 8)      # the warp-specialised row sweep forced (TMA, mbarriers)
    torch.cuda.synchronize()
    P.check_status()
    x = torch.randn(1, 2, H, W, device="cuda", requires_grad=True)
    P.geocyclic_pad(x, 2).sum().backward()
    print("case", H, W, interp, "ok", float(R.gfield.abs().sum()))

# GeoCyclic padding fused into the depthwise convolution (forward, grad input, grad weight / bias)
for (H, W, C, k) in [(33, 70, 3, 3), (40, 64, 2, 5), (17, 24, 2, 7)]:
    x = torch.randn(2, C, H, W, device="cuda", requires_grad=True)
    w = torch.randn(C, 1, k, k, device="cuda", requires_grad=True)
    b = torch.randn(C, device="cuda", requires_grad=True)
    P.geocyclic_dwconv(x, w, b).square().sum().backward()
    torch.cuda.synchronize()
    print("dwconv", H, W, k, "ok", float(x.grad.abs().sum()), float(w.grad.abs().sum()))
    y = P.geocyclic_avgpool5(x, 2)
    y.sum().backward()
